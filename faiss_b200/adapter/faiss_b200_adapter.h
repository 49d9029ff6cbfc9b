// faiss_b200 adapter -- the backend behind the reference's OWN C++ interface.
//
// Each class derives from faiss::Index (faiss/Index.h:101-435) and forwards to the C ABI of libfaiss_b200.so
// (include/faiss_b200_c.h), so everything in Faiss that drives a `faiss::Index&` -- faiss::Clustering::train
// (faiss/Clustering.cpp:254-356), faiss::IndexShards / ThreadedIndex (faiss/IndexShards.cpp:197-264),
// ProductQuantizer::assign_index, IndexIVF's quantizer slot -- runs on the B200 kernels without a change at
// the call site.  index_cpu_to_b200 / index_b200_to_cpu are the cloner pair of faiss/gpu/GpuCloner.cpp:124-255
// (copyFrom / copyTo of GpuIndexFlat.cu:105-176, GpuIndexIVFFlat.cu:89-150, GpuIndexIVFPQ.cu:105-217; inverted
// lists are moved in the CPU ArrayInvertedLists byte format, so copyTo(copyFrom(x)) is byte-identical).
//
// Compiled against the reference's headers only (no reference source is copied); built by
// tests/adapter/build_adapter.py where /root/reference is available and exercised by tests/adapter/.
#pragma once

#include <faiss/Index.h>
#include <faiss/IndexFlat.h>
#include <faiss/IndexIVF.h>
#include <faiss/IndexIVFFlat.h>
#include <faiss/IndexIVFPQ.h>

#include <memory>

struct FaissStandardGpuResources_H;
struct FaissIndex_H;

namespace faiss_b200_adapter {

// faiss::gpu::StandardGpuResources' role: owns streams, temp memory and NCCL communicators
class B200Resources {
   public:
    B200Resources();
    ~B200Resources();
    B200Resources(const B200Resources&) = delete;
    B200Resources& operator=(const B200Resources&) = delete;
    FaissStandardGpuResources_H* handle() const {
        return h_;
    }
    void ncclInitAll(const std::vector<int>& devices);

   private:
    FaissStandardGpuResources_H* h_ = nullptr;
};

// common part: an opaque C handle + the forwarding of the faiss::Index virtuals
class B200Index : public faiss::Index {
   public:
    ~B200Index() override;
    void train(faiss::idx_t n, const float* x) override;
    void add(faiss::idx_t n, const float* x) override;
    void add_with_ids(faiss::idx_t n, const float* x, const faiss::idx_t* xids) override;
    void search(
            faiss::idx_t n,
            const float* x,
            faiss::idx_t k,
            float* distances,
            faiss::idx_t* labels,
            const faiss::SearchParameters* params = nullptr) const override;
    void reset() override;
    void reconstruct(faiss::idx_t key, float* recons) const override;
    void reconstruct_n(faiss::idx_t i0, faiss::idx_t ni, float* recons) const override;
    FaissIndex_H* handle() const {
        return h_;
    }
    int device() const {
        return device_;
    }

   protected:
    B200Index(int d, faiss::MetricType metric, int device) : faiss::Index(d, metric), device_(device) {}
    void sync_();
    FaissIndex_H* h_ = nullptr;
    int device_;
};

class B200IndexFlat : public B200Index { // faiss::gpu::GpuIndexFlat (faiss/gpu/GpuIndexFlat.h:43-141)
   public:
    // useFloat16: GpuIndexFlatConfig::useFloat16 (faiss/gpu/GpuIndexFlat.h:26-35) -- fp16 storage, queries rounded to fp16
    B200IndexFlat(B200Resources* res, int dims, faiss::MetricType metric, int device = 0, bool useFloat16 = false);
    B200IndexFlat(B200Resources* res, const faiss::IndexFlat* index, int device = 0, bool useFloat16 = false);
    void copyFrom(const faiss::IndexFlat* index);
    void copyTo(faiss::IndexFlat* index) const;
};

class B200IndexIVF : public B200Index { // faiss::gpu::GpuIndexIVF (faiss/gpu/GpuIndexIVF.h:40-167)
   public:
    size_t nlist;
    size_t nprobe = 1;
    void search(
            faiss::idx_t n,
            const float* x,
            faiss::idx_t k,
            float* distances,
            faiss::idx_t* labels,
            const faiss::SearchParameters* params = nullptr) const override;

   protected:
    B200IndexIVF(int d, faiss::MetricType metric, size_t nlist_, int device) : B200Index(d, metric, device), nlist(nlist_) {}
    void copyListsFrom_(const faiss::IndexIVF* index);
    void copyListsTo_(faiss::IndexIVF* index) const;
};

class B200IndexIVFFlat : public B200IndexIVF { // faiss/gpu/GpuIndexIVFFlat.h:24-119
   public:
    B200IndexIVFFlat(B200Resources* res, int dims, size_t nlist, faiss::MetricType metric, int device = 0);
    B200IndexIVFFlat(B200Resources* res, const faiss::IndexIVFFlat* index, int device = 0);
    void copyFrom(const faiss::IndexIVFFlat* index);
    void copyTo(faiss::IndexIVFFlat* index) const;
};

class B200IndexIVFPQ : public B200IndexIVF { // faiss/gpu/GpuIndexIVFPQ.h:56-181
   public:
    B200IndexIVFPQ(B200Resources* res, int dims, size_t nlist, size_t M, size_t nbits, faiss::MetricType metric, int device = 0);
    B200IndexIVFPQ(B200Resources* res, const faiss::IndexIVFPQ* index, int device = 0);
    void copyFrom(const faiss::IndexIVFPQ* index);
    void copyTo(faiss::IndexIVFPQ* index) const;
    size_t M, nbits;
};

// faiss::gpu::index_cpu_to_gpu / index_gpu_to_cpu (faiss/gpu/GpuCloner.cpp:124-255) for the three index types on the path
// the fields of GpuClonerOptions (faiss/gpu/GpuClonerOptions.h:17-56) that act on this path
struct B200ClonerOptions {
    bool useFloat16 = false; // Flat: fp16 storage
};
faiss::Index* index_cpu_to_b200(B200Resources* res, int device, const faiss::Index* index, const B200ClonerOptions* options = nullptr);
faiss::Index* index_b200_to_cpu(const faiss::Index* index);

} // namespace faiss_b200_adapter
