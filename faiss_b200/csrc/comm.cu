// faiss_b200 -- NCCL communicator ownership.  See comm.h.
#include "comm.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "resources.h"

namespace fb200 {

namespace {

void* openNccl(std::string& tried) {
    // RTLD_NOLOAD first: if the process already maps an NCCL (PyTorch's bundled copy has the soname
    // libnccl.so.2) use that one, never a second copy
    std::vector<std::string> names;
    if (const char* e = getenv("FB200_NCCL_LIB"))
        names.push_back(e);
    names.push_back("libnccl.so.2");
    names.push_back("libnccl.so");
    for (auto& n : names) {
        if (void* h = dlopen(n.c_str(), RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))
            return h;
    }
    for (auto& n : names) {
        if (void* h = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL))
            return h;
        tried += n + " (" + (dlerror() ? dlerror() : "?") + ") ";
    }
    return nullptr;
}

} // namespace

const NcclApi& NcclApi::get() {
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        memset(&api, 0, sizeof(api));
        std::string tried;
        void* h = openNccl(tried);
        if (!h) {
            err = "cannot load NCCL (set FB200_NCCL_LIB): " + tried;
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(h, n);
            if (!p)
                err += std::string("missing symbol ") + n + "; ";
            return p;
        };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    if (!err.empty())
        FB_THROW_FMT("NCCL unavailable: %s", err.c_str());
    return api;
}

Communicator::~Communicator() {
    if (comm_) {
        try {
            DeviceScope scope(device_);
            NcclApi::get().CommDestroy(comm_);
        } catch (...) {
        }
    }
}

void Communicator::allGatherBytes(const void* send, void* recv, size_t bytes, cudaStream_t stream) const {
    NCCL_VERIFY(NcclApi::get().AllGather(send, recv, bytes, ncclChar, comm_, stream));
}

void Communicator::allGatherPair(
        const float* sendF, float* recvF, size_t countF, const idx_t* sendI, idx_t* recvI, size_t countI, cudaStream_t stream) const {
    const NcclApi& a = NcclApi::get();
    NCCL_VERIFY(a.GroupStart());
    NCCL_VERIFY(a.AllGather(sendF, recvF, countF, ncclFloat32, comm_, stream));
    NCCL_VERIFY(a.AllGather(sendI, recvI, countI, ncclInt64, comm_, stream));
    NCCL_VERIFY(a.GroupEnd());
}

void Communicator::allReduceMax(float* buf, size_t count, cudaStream_t stream) const {
    NCCL_VERIFY(NcclApi::get().AllReduce(buf, buf, count, ncclFloat32, ncclMax, comm_, stream));
}

void Communicator::allReduceSum(float* buf, size_t count, cudaStream_t stream) const {
    NCCL_VERIFY(NcclApi::get().AllReduce(buf, buf, count, ncclFloat32, ncclSum, comm_, stream));
}

void Communicator::broadcastBytes(void* buf, size_t bytes, int root, cudaStream_t stream) const {
    NCCL_VERIFY(NcclApi::get().Broadcast(buf, buf, bytes, ncclChar, root, comm_, stream));
}

std::vector<int64_t> Communicator::allGatherHostI64(int64_t v, cudaStream_t stream) const {
    DeviceScope scope(device_);
    int64_t* d = nullptr;
    CUDA_VERIFY(cudaMalloc(&d, sizeof(int64_t) * (nranks_ + 1)));
    std::vector<int64_t> out(nranks_);
    try {
        CUDA_VERIFY(cudaMemcpyAsync(d + nranks_, &v, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
        NCCL_VERIFY(NcclApi::get().AllGather(d + nranks_, d, 1, ncclInt64, comm_, stream));
        CUDA_VERIFY(cudaMemcpyAsync(out.data(), d, sizeof(int64_t) * nranks_, cudaMemcpyDeviceToHost, stream));
        CUDA_VERIFY(cudaStreamSynchronize(stream));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
    return out;
}

std::array<char, NCCL_UNIQUE_ID_BYTES> Communicator::uniqueId() {
    ncclUniqueId id;
    NCCL_VERIFY(NcclApi::get().GetUniqueId(&id));
    std::array<char, NCCL_UNIQUE_ID_BYTES> out;
    memcpy(out.data(), id.internal, NCCL_UNIQUE_ID_BYTES);
    return out;
}

std::shared_ptr<Communicator> Communicator::initRank(int device, int nranks, int rank, const char* id128) {
    FB_THROW_IF_NOT_MSG(nranks >= 1 && rank >= 0 && rank < nranks, "invalid rank / world size");
    DeviceScope scope(device);
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t c = nullptr;
    NCCL_VERIFY(NcclApi::get().CommInitRank(&c, nranks, id, rank));
    return std::make_shared<Communicator>(c, rank, nranks, device);
}

std::vector<std::shared_ptr<Communicator>> Communicator::initAll(const std::vector<int>& devices) {
    FB_THROW_IF_NOT_MSG(!devices.empty(), "no devices");
    std::vector<ncclComm_t> comms(devices.size(), nullptr);
    NCCL_VERIFY(NcclApi::get().CommInitAll(comms.data(), (int)devices.size(), devices.data()));
    std::vector<std::shared_ptr<Communicator>> out;
    for (size_t i = 0; i < devices.size(); i++)
        out.push_back(std::make_shared<Communicator>(comms[i], (int)i, (int)devices.size(), devices[i]));
    return out;
}

} // namespace fb200
