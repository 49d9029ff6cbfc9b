// faiss_b200 -- NCCL communicator ownership (L0) for the sharded paths.
//
// The reference shards a database over the GPUs of one box with IndexShards: one worker thread per
// sub-index, per-shard D2H copies and a host heap merge (faiss/IndexShards.cpp:197-264,
// faiss/impl/ThreadedIndex-inl.h:119-194, faiss/utils/Heap.cpp:166-238).  Here the exchange is a
// collective over NVLink / NVSwitch: the resources object owns one NCCL communicator per device
// (SURVEY 7 step 1 / step 8), either created for all devices of this process at once
// (ncclCommInitAll: the in-process IndexShards fast path) or joined by rank (ncclCommInitRank: one
// process per GPU, launched by torchrun or any other launcher that can hand 128 bytes of unique id
// to every rank).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a process that already carries an NCCL
// (PyTorch bundles one) keeps exactly that copy, and single-GPU users never load it.
#pragma once

#include <nccl.h> // types and enums only; every function is called through NcclApi

#include <array>
#include <memory>
#include <vector>

#include "common.h"

namespace fb200 {

struct NcclApi {
    ncclResult_t (*GetVersion)(int*);
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    // throws FaissException when no NCCL library can be loaded
    static const NcclApi& get();
};

#define NCCL_VERIFY(X)                                                                                  \
    do {                                                                                                \
        ncclResult_t __r = (X);                                                                         \
        if (__r != ncclSuccess) {                                                                       \
            FB_THROW_FMT("NCCL error %d: %s (%s)", (int)__r, ::fb200::NcclApi::get().GetErrorString(__r), #X); \
        }                                                                                               \
    } while (0)

// One rank of a communicator, bound to one device.  All collectives are enqueued on the caller's
// stream (the resources' ordering stream), so they are ordered with the kernels around them.
class Communicator {
   public:
    Communicator(ncclComm_t comm, int rank, int nranks, int device) : comm_(comm), rank_(rank), nranks_(nranks), device_(device) {}
    ~Communicator();
    Communicator(const Communicator&) = delete;
    Communicator& operator=(const Communicator&) = delete;

    int rank() const {
        return rank_;
    }
    int size() const {
        return nranks_;
    }
    int device() const {
        return device_;
    }
    ncclComm_t raw() const {
        return comm_;
    }

    // recv = concatenation over ranks of `bytes` bytes from each rank's send
    void allGatherBytes(const void* send, void* recv, size_t bytes, cudaStream_t stream) const;
    // two all-gathers fused into one NCCL launch: [nranks][countF] floats and [nranks][countI] int64
    void allGatherPair(const float* sendF, float* recvF, size_t countF, const idx_t* sendI, idx_t* recvI, size_t countI, cudaStream_t stream) const;
    void allReduceMax(float* buf, size_t count, cudaStream_t stream) const; // in place
    void allReduceSum(float* buf, size_t count, cudaStream_t stream) const; // in place
    void broadcastBytes(void* buf, size_t bytes, int root, cudaStream_t stream) const;
    // host-side convenience (tiny messages, synchronises `stream`): every rank's value
    std::vector<int64_t> allGatherHostI64(int64_t v, cudaStream_t stream) const;

    static std::array<char, NCCL_UNIQUE_ID_BYTES> uniqueId();
    static std::shared_ptr<Communicator> initRank(int device, int nranks, int rank, const char* id128);
    // one communicator per device of this process (rank i = devices[i])
    static std::vector<std::shared_ptr<Communicator>> initAll(const std::vector<int>& devices);

   private:
    ncclComm_t comm_;
    int rank_, nranks_, device_;
};

} // namespace fb200
