"""TEST INFRASTRUCTURE: compiles the faiss::Index adapter (faiss_b200/adapter) together with its driver
(tests/adapter/adapter_test.cpp) against the REFERENCE's headers (/root/reference) and links the reference CPU library
(oracle/_ref) -- the driver runs faiss::Clustering / faiss::IndexShards / the cloner pair over the adapter.  Only
possible where /root/reference exists; the binary (tests/adapter/_build/adapter_test, git-ignored) travels to the GPU
box with the snapshot."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.join(ROOT, "faiss_b200")
LIB = os.path.join(HERE, "libfaiss_b200.so")


def build_adapter(verbose=True):
    ref = "/root/reference"
    root = ROOT
    reflib = os.path.join(root, "oracle", "_ref", "libfaiss_ref.so")
    if not os.path.isdir(os.path.join(ref, "faiss")) or not os.path.exists(reflib):
        return None
    outdir = os.path.join(root, "tests", "adapter", "_build")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "adapter_test")
    srcs = [os.path.join(HERE, "adapter", "faiss_b200_adapter.cpp"), os.path.join(root, "tests", "adapter", "adapter_test.cpp")]
    deps = srcs + [os.path.join(HERE, "adapter", "faiss_b200_adapter.h"), LIB, reflib, os.path.join(root, "include", "faiss_b200_c.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(f) for f in deps):
        return out
    import sysconfig

    blasdir = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")  # OpenBLAS + libgfortran of libfaiss_ref
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-fopenmp", "-w", "-Wl,-rpath-link," + blasdir, "-I" + ref, "-I" + os.path.join(root, "include"),
           "-I" + os.path.join(HERE, "adapter")] + srcs + [
        "-o", out, reflib, LIB, "-L/usr/local/cuda/lib64", "-lcudart",
        "-Wl,-rpath,$ORIGIN/../../../oracle/_ref:$ORIGIN/../../../faiss_b200:/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("adapter build failed:\n%s\n%s" % (r.stdout[-3000:], r.stderr[-3000:]))
    if verbose:
        print("[tests.adapter] built", out, flush=True)
    return out



if __name__ == "__main__":
    print(build_adapter())
