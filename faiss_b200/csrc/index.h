// faiss_b200 -- host-side index objects (L2/L3): the drop-in surface.
//
// Mirrors the reference classes (same member names, argument meaning, limits and error behaviour):
//   faiss::Index                      faiss/Index.h:101-435
//   faiss::gpu::GpuIndex              faiss/gpu/GpuIndex.h:53-297, GpuIndex.cu
//   faiss::gpu::GpuIndexFlat{,L2,IP}  faiss/gpu/GpuIndexFlat.h:26-217, GpuIndexFlat.cu:28-457
//   faiss::gpu::GpuIndexIVF           faiss/gpu/GpuIndexIVF.h:40-167, GpuIndexIVF.cu
//   faiss::gpu::GpuIndexIVFFlat       faiss/gpu/GpuIndexIVFFlat.h:24-119
//   faiss::gpu::GpuIndexIVFPQ         faiss/gpu/GpuIndexIVFPQ.h:25-181, GpuIndexIVFPQ.cu:29-622
//   faiss::Clustering                 faiss/Clustering.h:22-229, Clustering.cpp:60-380
//   faiss::IndexShards                faiss/IndexShards.h:21-106, IndexShards.cpp:87-264
// Everything device-side goes through kernels.h.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "resources.h"

namespace fb200 {

// ------------------------------------------------------------------------------------------
// growable device array backed by GpuResources (role of DeviceVector, faiss/gpu/utils/DeviceVector.cuh)
// ------------------------------------------------------------------------------------------
template <typename T>
class DeviceVector {
   public:
    DeviceVector(GpuResources* res, int device, AllocType type) : res_(res), device_(device), type_(type) {}
    ~DeviceVector() {
        clear();
    }
    DeviceVector(const DeviceVector&) = delete;
    DeviceVector& operator=(const DeviceVector&) = delete;

    T* data() const {
        return data_;
    }
    size_t size() const {
        return size_;
    }
    size_t capacity() const {
        return cap_;
    }
    void clear() {
        if (data_)
            res_->deallocMemory(device_, data_);
        data_ = nullptr;
        size_ = cap_ = 0;
    }
    // ensure capacity >= n (exact if `exact`, else geometric growth), preserving contents
    void reserve(size_t n, cudaStream_t stream, bool exact = false) {
        if (n <= cap_)
            return;
        size_t ncap = exact ? n : std::max(n, cap_ + cap_ / 2);
        AllocRequest r;
        r.type = type_;
        r.device = device_;
        r.space = MemorySpace::Device;
        r.stream = stream;
        r.size = ncap * sizeof(T);
        T* nd = (T*)res_->allocMemory(r);
        if (size_ > 0) {
            CUDA_VERIFY(cudaMemcpyAsync(nd, data_, size_ * sizeof(T), cudaMemcpyDeviceToDevice, stream));
            CUDA_VERIFY(cudaStreamSynchronize(stream));
        }
        if (data_)
            res_->deallocMemory(device_, data_);
        data_ = nd;
        cap_ = ncap;
    }
    void resize(size_t n, cudaStream_t stream) {
        reserve(n, stream);
        size_ = n;
    }
    // append n elements from a host or device pointer
    void append(const T* src, size_t n, cudaStream_t stream) {
        if (n == 0)
            return;
        reserve(size_ + n, stream);
        CUDA_VERIFY(cudaMemcpyAsync(data_ + size_, src, n * sizeof(T), cudaMemcpyDefault, stream));
        size_ += n;
    }

   private:
    GpuResources* res_;
    int device_;
    AllocType type_;
    T* data_ = nullptr;
    size_t size_ = 0, cap_ = 0;
};

// ------------------------------------------------------------------------------------------
// faiss::Index
// ------------------------------------------------------------------------------------------
struct Index {
    int d;
    idx_t ntotal = 0;
    bool verbose = false;
    bool is_trained = true;
    MetricType metric_type;
    float metric_arg = 0.f;

    explicit Index(int d_ = 0, MetricType m = METRIC_L2) : d(d_), metric_type(m) {}
    virtual ~Index() = default;

    virtual void train(idx_t /*n*/, const float* /*x*/) {}
    virtual void add(idx_t n, const float* x) = 0;
    virtual void add_with_ids(idx_t n, const float* x, const idx_t* xids);
    virtual void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const = 0;
    virtual void assign(idx_t n, const float* x, idx_t* labels, idx_t k = 1) const;
    virtual void reset() = 0;
    virtual void reconstruct(idx_t key, float* recons) const;
    virtual void reconstruct_batch(idx_t n, const idx_t* keys, float* recons) const;
    virtual void reconstruct_n(idx_t i0, idx_t ni, float* recons) const;
    virtual void compute_residual(const float* x, float* residual, idx_t key) const;
    virtual void compute_residual_n(idx_t n, const float* xs, float* residuals, const idx_t* keys) const;
};

// faiss::SearchParameters / SearchParametersIVF (faiss/Index.h:88-93, faiss/IndexIVF.h:68-90): per-call overrides.
// GPU indexes accept them like the reference's (faiss/gpu/GpuIndexIVF.cu:383-406): no IDSelector, max_codes == 0.
struct SearchParameters {
    void* sel = nullptr; // IDSelector: not supported on the GPU path (must stay null, as upstream)
    virtual ~SearchParameters() {}
};
struct SearchParametersIVF : SearchParameters {
    size_t nprobe = 1;
    size_t max_codes = 0;
    SearchParameters* quantizer_params = nullptr;
};

// faiss::InterruptCallback (faiss/impl/AuxIndexStructures.h): a process-wide hook polled between query pages,
// add pages, search rounds and clustering iterations (the reference polls between tiles,
// faiss/gpu/impl/Distance.cu:245,266,403-405); when it returns true the running call throws
// "computation interrupted".
struct InterruptCallback {
    typedef int (*Fn)(void* ctx);
    static void set(Fn fn, void* ctx);
    static void clear();
    static bool is_interrupted();
    static void check(); // throws FaissException if the callback asks to stop
};

// ------------------------------------------------------------------------------------------
// GpuIndex
// ------------------------------------------------------------------------------------------
struct GpuIndexConfig { // faiss/gpu/GpuIndex.h:32-47
    int device = 0;
    MemorySpace memorySpace = MemorySpace::Device;
};

class GpuIndex : public Index {
   public:
    GpuIndex(std::shared_ptr<GpuResources> resources, int dims, MetricType metric, float metricArg, GpuIndexConfig config);

    int getDevice() const {
        return config_.device;
    }
    std::shared_ptr<GpuResources> getResources() {
        return resources_;
    }
    void setMinPagingSize(size_t size) {
        minPagedSize_ = size;
    }
    size_t getMinPagingSize() const {
        return minPagedSize_;
    }

    // x / ids / distances / labels may live on the host or on any device
    void add(idx_t n, const float* x) override;
    void add_with_ids(idx_t n, const float* x, const idx_t* ids) override;
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override;
    // faiss::Index::search(..., const SearchParameters* params) (faiss/Index.h:207-214, faiss/gpu/GpuIndex.cu:373-448)
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels, const SearchParameters* params) const;
    void assign(idx_t n, const float* x, idx_t* labels, idx_t k = 1) const override;
    void compute_residual(const float* x, float* residual, idx_t key) const override;
    void compute_residual_n(idx_t n, const float* xs, float* residuals, const idx_t* keys) const override;

   protected:
    // the per-call parameters of the search in flight (an index instance is not re-entrant, as upstream)
    mutable const SearchParameters* callParams_ = nullptr;
    virtual bool addImplRequiresIDs_() const = 0;
    virtual void addImpl_(idx_t n, const float* xDev, const idx_t* idsDev) = 0;
    virtual void searchImpl_(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev) const = 0;
    // host queries above the paging threshold: pinned double buffers + async-copy stream (GpuIndex.cu:620-788)
    bool searchFromCpuPaged_(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels, idx_t maxQ) const;

   public:
    // Shard-local half of a sharded search (device pointers, ids local to this shard).  `flatShard` is
    // non-null only when EVERY rank of the communicator runs the tensor-core Flat path: the ranks then pool
    // their thresholds after every round, and a shard may return fewer than k entries (-1 padded) -- those it
    // can prove are not in the global top-k.  All ranks must make the call with identical queries and k.
    virtual void searchShardDevice(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev, const FlatTcShard* flatShard) const {
        (void)flatShard;
        searchImpl_(n, xDev, k, dDev, iDev);
    }
    // would this index take the tensor-core Flat path for (its current size, k)?  (sharded search: the
    // pooled-threshold protocol is used only if this holds on every rank)
    virtual bool shardPoolingEligible(int /*k*/, idx_t /*n*/) const {
        return false;
    }

   protected:
    cudaStream_t stream_() const {
        return resources_->getDefaultStream(config_.device);
    }

    std::shared_ptr<GpuResources> resources_;
    GpuIndexConfig config_;
    size_t minPagedSize_ = size_t(256) << 20; // faiss/gpu/GpuIndex.cu kMinPageSize
};

// ------------------------------------------------------------------------------------------
// GpuIndexFlat
// ------------------------------------------------------------------------------------------
struct GpuIndexFlatConfig : GpuIndexConfig { // faiss/gpu/GpuIndexFlat.h:26-35
    bool useFloat16 = false;        // store the vectors as fp16; queries are rounded to fp16 too (FlatIndex.cu:112-136)
    bool useTensorCores = true;     // false: always take the exact SIMT kernel
};

class GpuIndexFlat : public GpuIndex {
   public:
    GpuIndexFlat(
            std::shared_ptr<GpuResources> resources,
            int dims,
            MetricType metric,
            GpuIndexFlatConfig config = GpuIndexFlatConfig());
    ~GpuIndexFlat() override;

    // copyFrom / copyTo against the CPU IndexFlat payload (xb, row-major fp32)
    // (faiss/gpu/GpuIndexFlat.cu:105-176: index->get_xb(), ntotal)
    void copyFrom(idx_t n, const float* xb);
    void copyTo(float* xbOut) const;
    size_t getNumVecs() const {
        return (size_t)ntotal;
    }

    void reset() override;
    void train(idx_t n, const float* x) override;
    void add(idx_t n, const float* x) override;
    void add_with_ids(idx_t n, const float* x, const idx_t* ids) override; // unsupported, as upstream
    void reconstruct(idx_t key, float* out) const override;
    void reconstruct_n(idx_t i0, idx_t num, float* out) const override;
    void reconstruct_batch(idx_t n, const idx_t* keys, float* out) const override;
    void compute_residual(const float* x, float* residual, idx_t key) const override;
    void compute_residual_n(idx_t n, const float* xs, float* residuals, const idx_t* keys) const override;

    // device-pointer entry points used by IVF / clustering (role of FlatIndex::query)
    void searchDevice(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev) const {
        searchImpl_(n, xDev, k, dDev, iDev);
    }
    void searchShardDevice(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev, const FlatTcShard* flatShard) const override;
    bool shardPoolingEligible(int k, idx_t n) const override;
    const float* vectorsDevice() const { // fp32 rows (IVF coarse centroids); not available under useFloat16
        FB_THROW_IF_NOT_MSG(!flatConfig_.useFloat16, "fp32 row access on a float16 GpuIndexFlat");
        return vecs_.data();
    }
    bool usesFloat16() const {
        return flatConfig_.useFloat16;
    }
    void setUseTensorCores(bool v) {
        flatConfig_.useTensorCores = v;
    }
    // replace the whole content by n device rows without giving the storage back (the k-means loop installs a new
    // centroid table every iteration: reset() + add() would free and re-allocate five buffers each time)
    void replaceVectorsDevice(idx_t n, const float* xDev);
    // diagnostics
    mutable int lastSearchUsedTensorCores = 0;
    mutable int lastSearchFallbackQueries = 0;

   protected:
    bool addImplRequiresIDs_() const override {
        return false;
    }
    void addImpl_(idx_t n, const float* xDev, const idx_t* idsDev) override;
    void searchImpl_(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev) const override;
    void prepareTensorCoreData_() const;

    const void* rows_() const { // the stored rows, as the kernels take them (with yHalf_())
        return flatConfig_.useFloat16 ? (const void*)vecs16_.data() : (const void*)vecs_.data();
    }
    int yHalf_() const {
        return flatConfig_.useFloat16 ? 1 : 0;
    }
    // useFloat16: the query block rounded to fp16 and widened again (what the reference's convertTensor does before
    // its half GEMM); returns x itself otherwise
    const float* roundedQueries_(idx_t n, const float* xDev, GpuMemoryReservation& hold) const;

    GpuIndexFlatConfig flatConfig_;
    DeviceVector<float> vecs_;    // fp32 storage (default)
    DeviceVector<__half> vecs16_; // fp16 storage (useFloat16): the only copy of the vectors besides the scoring tiles
    // tensor-core side data, rebuilt lazily after adds
    mutable DeviceVector<__half> y16_;
    mutable DeviceVector<float> bias_;
    mutable DeviceVector<int> perm_;          // L2: stored (norm-sorted) position -> row id
    mutable DeviceVector<float> tileMaxBias_; // max bias per 256-row tile
    mutable bool tcDirty_ = true;
    mutable float yScale_ = 1.f;
    mutable float yMaxNorm_ = 0.f;
    mutable int dpad_ = 0;
};

class GpuIndexFlatL2 : public GpuIndexFlat {
   public:
    GpuIndexFlatL2(std::shared_ptr<GpuResources> r, int dims, GpuIndexFlatConfig c = GpuIndexFlatConfig())
            : GpuIndexFlat(std::move(r), dims, METRIC_L2, c) {}
};
class GpuIndexFlatIP : public GpuIndexFlat {
   public:
    GpuIndexFlatIP(std::shared_ptr<GpuResources> r, int dims, GpuIndexFlatConfig c = GpuIndexFlatConfig())
            : GpuIndexFlat(std::move(r), dims, METRIC_INNER_PRODUCT, c) {}
};

// ------------------------------------------------------------------------------------------
// Clustering (Lloyd k-means, training set resident on the device)
// ------------------------------------------------------------------------------------------
struct ClusteringParameters { // faiss/Clustering.h:22-78
    int niter = 25;
    int nredo = 1;
    bool verbose = false;
    bool spherical = false;
    bool int_centroids = false;
    bool update_index = false;
    bool frozen_centroids = false;
    int min_points_per_centroid = 39;
    int max_points_per_centroid = 256;
    int seed = 1234;
};

struct ClusteringIterationStats { // faiss/Clustering.h:80-92
    float obj;
    double time;
    double time_search;
    double imbalance_factor;
    int nsplit;
};

struct Clustering : ClusteringParameters {
    size_t d;
    size_t k;
    std::vector<float> centroids; // (k * d), host copy, as upstream
    std::vector<ClusteringIterationStats> iteration_stats;

    Clustering(int d, int k) : d(d), k(k) {}
    Clustering(int d, int k, const ClusteringParameters& cp) : ClusteringParameters(cp), d(d), k(k) {}
    // x: host or device pointer.  `index` is the assignment index (reset / add / search k=1).
    void train(idx_t n, const float* x, GpuIndexFlat& index);
    // The same Lloyd iterations with the training set SHARDED over the ranks of `comm` (rank order = row order of
    // the concatenated set) and the centroid table replicated: local Flat k=1 assignment, local partial sums
    // (deterministic sort + segmented sum), ONE packed ncclAllReduce per iteration (k*d sums | k counts | objective),
    // empty clusters refilled by the reference's deterministic split_clusters on every rank identically
    // (SURVEY 8(e)).  Collective: every rank calls it with its own rows; sub-sampling is the caller's business.
    // Every rank ends with identical centroids (also added to `index`).  splitSeconds: host time inside split_clusters.
    void trainSharded(idx_t nLocal, const float* xLocal, GpuIndexFlat& index, const Communicator& comm);
    double splitSeconds = 0;
};

// ProductQuantizer::train, Train_default (faiss/impl/ProductQuantizer.cpp:130-195); x device [n,d], pqOut host [M][256][d/M]
void trainProductQuantizer(
        std::shared_ptr<GpuResources> resources,
        int device,
        idx_t n,
        const float* xDev,
        int d,
        int M,
        const ClusteringParameters& cp,
        float* pqOut);

// helpers restated from the reference so that seeds reproduce its sampling decisions
void rand_perm(int* perm, size_t n, int64_t seed);                   // faiss/utils/random.cpp:188-199
int split_clusters(size_t d, size_t k, size_t n, float* hassign, float* centroids); // ClusteringHelpers.cpp:177-240

// ------------------------------------------------------------------------------------------
// inverted-list storage
// ------------------------------------------------------------------------------------------
class IvfLists {
   public:
    // pqInterleaved: store codes in the rotated, interleaved-by-32 PQ layout (kernels.h); the
    // host-facing accessors below always speak the flat [len][codeSize] ArrayInvertedLists format
    IvfLists(GpuResources* res, int device, int64_t nlist, int codeSize, bool pqInterleaved = false);
    ~IvfLists();
    void reset();
    void reserve(size_t totalVecs, cudaStream_t stream);
    // exact per-list capacities in ONE relayout (bulk copyFrom: all list lengths are known up front)
    void reserveLists(const int64_t* lens, cudaStream_t stream);
    // append n encoded rows (device pointers); assign[i] in [0,nlist) or -1 (skipped)
    // returns the number of rows actually stored
    idx_t append(idx_t n, const uint8_t* rowsDev, const idx_t* idsDev, const idx_t* assignDev, cudaStream_t stream);
    // bulk load of one list from host memory (copyFrom)
    void setListFromHost(int64_t l, int64_t len, const uint8_t* codes, const idx_t* ids, cudaStream_t stream);
    void getListToHost(int64_t l, uint8_t* codes, idx_t* ids, cudaStream_t stream) const;
    int64_t listLength(int64_t l) const {
        return hLen_[l];
    }
    size_t reclaim(cudaStream_t stream);

    const int64_t* dStart() const {
        return dStart_;
    }
    const int* dLen() const {
        return dLen_;
    }
    const uint8_t* codes() const {
        return codes_;
    }
    const idx_t* ids() const {
        return ids_;
    }
    int64_t nlist() const {
        return nlist_;
    }
    int codeSize() const {
        return codeSize_;
    }
    int maxListLength() const;
    int64_t arenaElems() const {
        return arenaElems_;
    }
    bool interleaved() const {
        return interleaved_;
    }

   private:
    void relayout_(const std::vector<int64_t>& newCap, cudaStream_t stream);
    void uploadMeta_(cudaStream_t stream);

    GpuResources* res_;
    int device_;
    int64_t nlist_;
    int codeSize_;
    bool interleaved_ = false;
    uint8_t* codes_ = nullptr;
    idx_t* ids_ = nullptr;
    int64_t arenaElems_ = 0;
    std::vector<int64_t> hStart_, hCap_;
    std::vector<int> hLen_;
    int64_t* dStart_ = nullptr;
    int* dLen_ = nullptr;
    int* dCounts_ = nullptr;
};

// ------------------------------------------------------------------------------------------
// GpuIndexIVF
// ------------------------------------------------------------------------------------------
enum IndicesOptions { INDICES_CPU = 0, INDICES_IVF = 1, INDICES_32_BIT = 2, INDICES_64_BIT = 3 };

struct GpuIndexIVFConfig : GpuIndexConfig { // faiss/gpu/GpuIndexIVF.h:23-35
    IndicesOptions indicesOptions = INDICES_64_BIT;
    GpuIndexFlatConfig flatConfig;
    bool allowCpuCoarseQuantizer = false;
};

class GpuIndexIVF : public GpuIndex {
   public:
    GpuIndexIVF(
            std::shared_ptr<GpuResources> resources,
            int dims,
            MetricType metric,
            idx_t nlist,
            int codeSize,
            GpuIndexIVFConfig config,
            bool pqInterleaved = false);
    ~GpuIndexIVF() override;

    idx_t nlist;
    size_t nprobe = 1;
    size_t max_codes = 0;
    GpuIndexFlat* quantizer = nullptr;
    // bumped whenever the coarse centroids are (re)installed through this class (train, setCoarseCentroids);
    // derived classes key centroid-dependent side tables on it
    uint64_t coarseEpoch = 0;
    bool own_fields = true;
    ClusteringParameters cp;

    idx_t getNumLists() const {
        return nlist;
    }
    idx_t getListLength(idx_t listId) const;
    std::vector<uint8_t> getListVectorData(idx_t listId) const;
    std::vector<idx_t> getListIndices(idx_t listId) const;
    void reserveMemory(size_t numVecs);
    // bulk-clone helper: exact capacity for every list in one arena relayout, before nlist x setList
    // (the role of the per-list reserve in IVFBase::copyInvertedListsFrom, faiss/gpu/impl/IVFBase.cu:328-451)
    void setListSizes(const idx_t* lens);
    size_t reclaimMemory();
    void reset() override;
    // share an existing coarse quantiser instead of the internally created one (not owned)
    void setQuantizer(GpuIndexFlat* coarse);
    // install coarse centroids [nlist,d] (copyFrom of the CPU quantizer's xb)
    void setCoarseCentroids(const float* centroidsHostOrDev);
    void getCoarseCentroids(float* out) const;
    // bulk list load (copyFrom ArrayInvertedLists): codes [len*code_size], ids [len]
    void setList(idx_t listId, idx_t len, const uint8_t* codes, const idx_t* ids);

    // faiss/gpu/GpuIndexIVF.cu:408-488
    void search_preassigned(
            idx_t n,
            const float* x,
            idx_t k,
            const idx_t* assign,
            const float* centroid_dis,
            float* distances,
            idx_t* labels) const;

   protected:
    bool addImplRequiresIDs_() const override {
        return true;
    }
    void trainQuantizer_(idx_t n, const float* xDev);
    // is a trained coarse quantiser all the training this index needs? (IVF-Flat: yes; IVF-PQ: the PQ too)
    virtual bool quantizerOnlyTraining_() const {
        return true;
    }
    void searchImpl_(idx_t n, const float* xDev, int k, float* dDev, idx_t* iDev) const override;
    virtual void scanImpl_(
            idx_t n,
            const float* xDev,
            const idx_t* probesDev,
            const float* coarseDisDev,
            int nprobe,
            int k,
            float* dDev,
            idx_t* iDev) const = 0;

    GpuIndexIVFConfig ivfConfig_;
    std::unique_ptr<IvfLists> lists_;
};

class GpuIndexIVFFlat : public GpuIndexIVF {
   public:
    GpuIndexIVFFlat(
            std::shared_ptr<GpuResources> resources,
            int dims,
            idx_t nlist,
            MetricType metric = METRIC_L2,
            GpuIndexIVFConfig config = GpuIndexIVFConfig());
    // faiss/gpu/GpuIndexIVFFlat.h:48-59: with an external (shared) coarse quantiser
    GpuIndexIVFFlat(
            std::shared_ptr<GpuResources> resources,
            GpuIndexFlat* coarseQuantizer,
            int dims,
            idx_t nlist,
            MetricType metric = METRIC_L2,
            GpuIndexIVFConfig config = GpuIndexIVFConfig());
    void train(idx_t n, const float* x) override;

   protected:
    void addImpl_(idx_t n, const float* xDev, const idx_t* idsDev) override;
    void scanImpl_(idx_t, const float*, const idx_t*, const float*, int, int, float*, idx_t*) const override;
};

struct GpuIndexIVFPQConfig : GpuIndexIVFConfig { // faiss/gpu/GpuIndexIVFPQ.h:25-49
    bool useFloat16LookupTables = false;
    bool usePrecomputedTables = false;
    bool interleavedLayout = false;
    bool useMMCodeDistance = false;
};

class GpuIndexIVFPQ : public GpuIndexIVF {
   public:
    GpuIndexIVFPQ(
            std::shared_ptr<GpuResources> resources,
            int dims,
            idx_t nlist,
            idx_t subQuantizers,
            idx_t bitsPerCode,
            MetricType metric = METRIC_L2,
            GpuIndexIVFPQConfig config = GpuIndexIVFPQConfig());
    // faiss/gpu/GpuIndexIVFPQ.h:69-82: with an external (shared) coarse quantiser
    GpuIndexIVFPQ(
            std::shared_ptr<GpuResources> resources,
            GpuIndexFlat* coarseQuantizer,
            int dims,
            idx_t nlist,
            idx_t subQuantizers,
            idx_t bitsPerCode,
            MetricType metric = METRIC_L2,
            GpuIndexIVFPQConfig config = GpuIndexIVFPQConfig());
    ~GpuIndexIVFPQ() override;

    int getNumSubQuantizers() const {
        return M_;
    }
    int getBitsPerCode() const {
        return nbits_;
    }
    int getCentroidsPerSubQuantizer() const {
        return 1 << nbits_;
    }
    // Precomputed term-2 table  T2[list][code][m] = ||y||^2 + 2 <centroid_list|m, y>  (the role of
    // faiss/gpu/impl/IVFPQ.cu:362-489 and IndexIVFPQ::precompute_table, faiss/IndexIVFPQ.cpp:376-458).
    // With it the per-(query, list) lookup table is T2[list] + (-2 <x|m, y>) -- one load and one add per
    // entry instead of dsub multiply-adds, which is what matters when lists are short (nlist = 65536).
    // L2 + interleaved layout only.  Policy: on when explicitly enabled; otherwise "auto" = the table is at
    // most precomputed_table_max_bytes (2 GiB, faiss/IndexIVFPQ.cpp:345) AND lists are short (< 4096 vectors
    // on average) -- with long lists the direct on-chip build is faster (see GpuIndexIVFPQ::precomputedActive_).
    void setPrecomputedCodes(bool enable) {
        usePrecomputed_ = enable;
        precomputedExplicit_ = true;
    }
    bool getPrecomputedCodes() const {
        return usePrecomputed_;
    }
    ClusteringParameters pq_cp; // ProductQuantizer::cp (faiss/impl/ProductQuantizer.h)
    void train(idx_t n, const float* x) override;
    // PQ centroids, layout [M][ksub][dsub] as in faiss::ProductQuantizer::centroids
    void setPQCentroids(const float* c);
    void getPQCentroids(float* out) const;

   protected:
    bool quantizerOnlyTraining_() const override {
        return false;
    }
    void trainResidualQuantizer_(idx_t n, const float* xDev);
    void addImpl_(idx_t n, const float* xDev, const idx_t* idsDev) override;
    void scanImpl_(idx_t, const float*, const idx_t*, const float*, int, int, float*, idx_t*) const override;

    int M_, nbits_;
    bool precomputedActive_() const;
    void ensureTerm2_() const;

    bool usePrecomputed_ = false;
    bool precomputedExplicit_ = false;
    uint64_t pqEpoch_ = 0;
    mutable DeviceVector<float> term2_; // [nlist][256][M]
    mutable uint64_t term2Key_ = ~uint64_t(0);
    DeviceVector<float> pqCentroids_;  // [M][256][dsub]
    DeviceVector<float> pqCentroidsT_; // [256][M][dsub] (LUT build reads it coalesced)
};

// ------------------------------------------------------------------------------------------
// IndexShards
// ------------------------------------------------------------------------------------------
class DistributedIndexShards;

class IndexShards : public Index {
   public:
    explicit IndexShards(int d, bool threaded = false, bool successive_ids = true);
    ~IndexShards() override;
    bool own_indices = false;
    bool threaded;
    bool successive_ids;

    void add_shard(Index* idx);
    void remove_shard(Index* idx);
    int count() const {
        return (int)shards_.size();
    }
    Index* at(int i) {
        return shards_[i];
    }
    void train(idx_t n, const float* x) override;
    void add(idx_t n, const float* x) override;
    void add_with_ids(idx_t n, const float* x, const idx_t* xids) override;
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override;
    void reset() override;
    void syncWithSubIndexes();
    // which path the last search took: 0 = thread-per-shard + host merge (the reference's), 1 = NCCL fast path
    mutable int lastSearchPath = 0;

   protected:
    template <typename F>
    void runOnIndex(F f) const;
    // fast path: every shard is a GpuIndex on its own device and the shards' resources hold one NCCL clique
    // over exactly those devices -> per-device threads + ncclAllGather + device merge (no host merge)
    bool ncclFastPath_(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const;
    std::vector<Index*> shards_;
    mutable std::vector<std::unique_ptr<DistributedIndexShards>> dist_; // one per device, built on first use
};

// faiss::IndexShardsIVF (faiss/IndexShardsIVF.h, IndexShardsIVF.cpp:100-251): shards are IVF indexes over ONE common
// coarse quantiser (GpuMultipleClonerOptions::common_ivf_quantizer, faiss/gpu/GpuCloner.cpp:418-436): the coarse search
// runs once, every shard scans its part of the probed lists through search_preassigned, results are merged.
class IndexShardsIVF : public IndexShards {
   public:
    IndexShardsIVF(GpuIndexFlat* quantizer, idx_t nlist, bool threaded = false, bool successive_ids = true);
    GpuIndexFlat* quantizer; // shared, not owned
    idx_t nlist;
    void add_shard(Index* idx); // must be a GpuIndexIVF with the same nlist
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override;
};

// ------------------------------------------------------------------------------------------
// IndexShards across NCCL ranks: one shard (a GpuIndex) per rank of the resources' communicator for the
// shard's device.  Same semantics as faiss::IndexShards::search (faiss/IndexShards.cpp:197-264): every query
// goes to every shard, ids are translated by the number of vectors in lower-ranked shards when successive_ids,
// results are merged with the (distance, id) rule of merge_knn_results (faiss/utils/Heap.cpp:166-238) --
// but the exchange is ONE grouped ncclAllGather of the per-shard [n,k] blocks over NVLink and the merge a
// device kernel reading the gathered layout in place.  With Flat shards on the tensor-core path the ranks
// also pool their k-th-score thresholds after every round (one small all-reduce), so per-query work
// (candidate selection, exact re-rank) shrinks with the number of shards instead of being replicated.
// Used by: one process per GPU (every process holds one instance; search() is a collective call), and by
// IndexShards' in-process fast path (one instance per device, driven by one thread per device).
// ------------------------------------------------------------------------------------------
class DistributedIndexShards : public Index {
   public:
    DistributedIndexShards(std::shared_ptr<GpuResources> resources, GpuIndex* local, bool successive_ids = true);
    ~DistributedIndexShards() override;
    bool own_local = false;
    bool successive_ids;

    int rank() const;
    int worldSize() const;
    idx_t idOffset() const {
        return idOffset_;
    }
    GpuIndex* local() {
        return local_;
    }
    // collective: re-reads every shard's ntotal (call after adds)
    void syncWithSubIndexes();
    void train(idx_t n, const float* x) override;           // local shard trains on x
    void add(idx_t n, const float* x) override;              // adds x to the LOCAL shard, then syncWithSubIndexes()
    void add_with_ids(idx_t n, const float* x, const idx_t* xids) override;
    void reset() override;
    // collective; x / distances / labels host or device, identical x on every rank
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override;
    // the collective itself; a rank that does not need the merged result (in-process fast path: only one
    // caller-visible output) skips the merge
    void searchCollective(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels, bool wantResult) const;
    idx_t shardSize(int r) const {
        return sizes_[r];
    }

   private:
    std::shared_ptr<GpuResources> resources_;
    GpuIndex* local_;
    std::shared_ptr<Communicator> comm_;
    std::vector<idx_t> sizes_;
    bool allFlatTc_ = false; // every rank's shard is a tensor-core-capable GpuIndexFlat
    idx_t idOffset_ = 0;
    int64_t maxTiles_ = 0;
    idx_t* dOffsets_ = nullptr; // device [world]
};

// host merge with the reference semantics (faiss/utils/Heap.cpp:166-238)
void merge_knn_results_host(
        idx_t n,
        idx_t k,
        int nshard,
        MetricType metric,
        const float* all_distances,
        const idx_t* all_labels,
        float* distances,
        idx_t* labels);

} // namespace fb200
