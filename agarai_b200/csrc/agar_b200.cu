// agar_b200.cu — C-ABI of the engine (include/agar_b200.h): handle management, launches, and the
// host-buffer convenience call.  Built for sm_100a only (see agarai_b200/build.py).
#include <cuda_runtime.h>

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/agar_b200.h"
#include "../../include/agar_spec.h"
#include "agar_host_expand.h"
#include "agar_kernels.cuh"
#include "agar_features.cuh"

using namespace agar;

namespace {

thread_local std::string g_err;

int fail(const std::string& m) { g_err = m; return 1; }

#define CUDA_OK(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
    } while (0)

constexpr int kThreads = 128;

using StepKernel = void (*)(DevParams, KArgs);

// Compile-time specialisations of the step kernel for the BASELINE worlds (agar_kernels.cuh, "config views"):
// <agents, bots, pellets, viruses, food slots, grid, channel flags, ticks>, each with the CTAs-per-SM bound its
// shared-memory footprint allows on B200.  Any other config runs the RtCfg instantiation.
using CvCfg1 = StCfg<1, 0, 1000, 0, 64, 64, 7, 4>;    // configs[0]/[2]/[4]: single agent, pellets only
using CvCfg2 = StCfg<1, 25, 1000, 25, 64, 64, 7, 4>;  // configs[1]: 25 bots + 25 viruses
using CvCfg4 = StCfg<16, 0, 1000, 0, 64, 64, 7, 4>;   // configs[3]: 16 agents per arena
// the reference's own default worlds
using CvA2C = StCfg<32, 0, 1000, 25, 64, 32, 15, 4>;     // a2c/hyperparameters.py:89-107: 32 agents, 25 viruses, grid 32
using CvA2CTest = StCfg<8, 25, 1000, 25, 64, 32, 15, 4>; // a2c/train.py:17-34: 8 agents + 25 bots
using CvDQN = StCfg<1, 0, 30, 0, 64, 128, 15, 4>;        // dqn/hyperparameters.py:23-28,97-104: arena 45, 30 pellets, grid 128
#ifdef AGAR_EXTRA_CV
// one more world, given on the compiler command line: agarai_b200.build.build_specialised(cfg) compiles a copy of
// the library with -DAGAR_EXTRA_CV -DAGAR_EXTRA_A=.. -DAGAR_EXTRA_BOTS=.. ... for a config outside the list above
using CvExtra = StCfg<AGAR_EXTRA_A, AGAR_EXTRA_BOTS, AGAR_EXTRA_P, AGAR_EXTRA_V, AGAR_EXTRA_F, AGAR_EXTRA_G, AGAR_EXTRA_FLAGS, AGAR_EXTRA_TICKS>;
#endif

template <class CV>
bool matches(const AgarConfig& c, const AgarLayout& L) {
    typedef typename CV::Layout SL;
    return c.num_agents == CV::A && c.num_bots == CV::BOTS && c.num_pellets == CV::P && c.num_viruses == CV::V &&
           c.max_foods == CV::F && c.grid_size == CV::G && c.obs_flags == CV::FLAGS && c.ticks_per_step == CV::TICKS &&
           c.event_capacity == 0 && L.words == SL::words && L.cell_m == SL::cell_m && L.food_vy == SL::food_vy &&
           L.agent_done == SL::agent_done && L.tick == SL::tick && L.dyn_begin == SL::dyn_begin;
}

// Host worker threads of agar_step_host's compact copy: fn(worker, workers) runs on every worker.
class HostPool {
public:
    explicit HostPool(int n) : n_(n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); });
        pin();
    }
    // Each worker stays on one core of this rank's share of the process's CPU set (AGAR_HOST_PIN=0 leaves them to the
    // scheduler): worker w updates the same agents every step, and the lines it wrote in the previous step - the bins it
    // now clears - are still in that core's L2; ranks sharing a host keep out of each other's cores.
    void pin() {
        if (const char* v = std::getenv("AGAR_HOST_PIN")) if (std::atoi(v) == 0) return;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) != 0) return;
        std::vector<int> cpus;
        for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) cpus.push_back(c);
        int ranks = 1, rank = 0;
        if (const char* v = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::atoi(v) > 0 ? std::atoi(v) : 1;
        if (const char* v = std::getenv("LOCAL_RANK")) rank = std::atoi(v);
        const int per = (int)cpus.size() / ranks;
        if (per < 2 || rank < 0 || rank >= ranks || n_ >= per) return;
        caller_cpu_ = cpus[(size_t)rank * per];
        for (int i = 0; i < n_; ++i) {  // the share's first core is left to the calling thread (CallerPin)
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[(size_t)rank * per + (size_t)((i + 1) % per)], &one);
            pthread_setaffinity_np(th_[(size_t)i].native_handle(), sizeof(one), &one);
        }
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // start(fn) hands fn to every worker and returns; wait() returns when all of them are done (fn must live until then).
    // Neither side sleeps while a call is in flight: a sleeping vCPU of this kind of host takes 50-150 us to come back
    // (measured as a bimodal tail of the call).  A worker polls for the next job for a grace period after finishing one
    // (back-to-back calls find it awake), then sleeps on the condition variable; wait() polls the pending counter.
    void start(const std::function<void(int, int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &fn;
            pending_.store(n_, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    }
    void wait() {
        for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins)
            if ((spins & 4095) == 4095) std::this_thread::yield(); else _mm_pause();
    }
    int caller_cpu() const { return caller_cpu_; }
    int size() const { return n_; }

private:
    void loop(int i) {
        int seen = 0;
        for (;;) {
            // poll for ~200 us, then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int spins = 0;; ++spins) {
                if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
                _mm_pause();
                if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(m_);
                sleepers_.fetch_add(1, std::memory_order_acq_rel);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_acq_rel);
            }
            const std::function<void(int, int)>* f;
            {
                std::lock_guard<std::mutex> lk(m_);  // (job_ and stop_ are written under the lock, before gen_ moves)
                if (stop_) return;
                seen = gen_.load(std::memory_order_acquire);
                f = job_;
            }
            (*f)(i, n_);
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    int caller_cpu_ = -1;  // >= 0 when the workers are pinned: the core of this rank's share they leave free
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int, int)>* job_ = nullptr;
    std::atomic<int> gen_{0}, pending_{0}, sleepers_{0};
    bool stop_ = false;
};

// While agar_step_host polls the chunk events its thread stays on the core the pinned workers leave free, so that it
// never shares a core with one of them; the caller's affinity mask is restored on return.
struct CallerPin {
    cpu_set_t saved;
    bool active = false;
    explicit CallerPin(int cpu) {
        if (cpu < 0 || sched_getaffinity(0, sizeof(saved), &saved) != 0) return;
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpu, &one);
        active = sched_setaffinity(0, sizeof(one), &one) == 0;
    }
    ~CallerPin() { if (active) sched_setaffinity(0, sizeof(saved), &saved); }
};

constexpr int kMaxWorkers = 32;  // (host_threads_default caps the pool at 32)
struct alignas(64) TakeCounter { std::atomic<int> v{0}; };  // one cache line each: owners and thieves of different lists do not collide
constexpr int kHostChunks = 4;   // agar_step_host steps the batch in chunks: copies and host work overlap the next kernel
constexpr int kEntryCap = 256;   // (index, value) pairs kept per agent image; a fuller image is copied whole

}  // namespace

struct AgarEnv {
    DevParams p;
    int device = 0;
    float* d_state = nullptr;
    int32_t* d_events = nullptr;
    int32_t* d_ev_counts = nullptr;
    int32_t* d_hash = nullptr;  // persisted pellet hash, one record per arena (derived data, not state)
    float* d_tab_xy = nullptr;
    int32_t* d_tab_act = nullptr;
    int32_t n_actions = 0;
    size_t smem = 0;
    int min_blocks = 5;  // which agar_step_kernel instantiation to launch
    int threads = kThreads;  // threads per arena (CTA size) of that instantiation
    StepKernel kernel = nullptr;
    const char* kernel_name = "";
    int64_t launches = 0;
    // costly-arenas-first launch order (KArgs::order_*): three order buffers of 8 + 4 * N words (read / written / cleared)
    int32_t* d_order = nullptr;
    int order_cur = -1;        // which list holds a complete order (-1: none yet)
    bool order_enabled = true; // AGAR_NO_ORDER (debug switch) read once at create
    // staging for agar_step_host
    float* d_target = nullptr; int32_t* d_act = nullptr; float* d_obs = nullptr;
    float* d_reward = nullptr; uint8_t* d_done = nullptr;
    cudaStream_t host_stream = nullptr, host_stream2 = nullptr;
    cudaEvent_t host_event = nullptr;
    // ordering of agar_step_host (private streams) against work the caller enqueued on its own streams
    cudaEvent_t order_event = nullptr;
    cudaStream_t last_stream = nullptr;
    bool pending = false;        // work was enqueued on last_stream since the last agar_step_host
    bool pending_multi = false;  // ... on more than one caller stream
    bool persist_hash = true;    // AGAR_NO_PERSIST (debug switch) read once at create
    // compact device-to-host copy of agar_step_host (obs_compact_kernel): occupied bins + centroids instead of images
    bool host_compact = false;
    int entry_cap = 0;
    ObsEntry* d_entries = nullptr; float* d_meta = nullptr;  // meta: per agent {centroid x, y, has cells, list length (int32 bits)}
    ObsEntry* h_entries = nullptr; float* h_meta = nullptr;  // pinned, two halves: this step's / the previous step's
    // persistent host observation buffer (agar_host_obs_persistent): the previous step's lists describe what h_obs holds
    bool host_persistent = false;
    int hbuf = 0;                      // which half of h_entries / h_meta the last completed call filled
    const float* delivered = nullptr;  // h_obs of the last completed compact call (nullptr: contents unknown)
    // AGAR_HOST_PROFILE (tuning aid, read once at create): where agar_step_host's wall time goes, printed by agar_destroy
    bool host_profile = false;
    double prof_us[4] = {0, 0, 0, 0};  // enqueue | waiting for a chunk's lists | host threads | final stream syncs
    std::atomic<long long> prof_tsc[5] = {};  // summed over the workers: waiting for a chunk | plan | prefetch | in-place update | full rewrite
    int64_t prof_calls = 0;
    TakeCounter* take = nullptr;        // [kHostChunks][kMaxWorkers] agents taken from each worker's list of each chunk (this call)
    long long prof_fin[64] = {};       // TSC at which each worker finished the call's job
    double prof_spread[3] = {0, 0, 0};  // kilo-cycles from the last chunk's publication to the first / median / last worker's finish
    cudaEvent_t chunk_done[kHostChunks] = {};
    HostPool* pool = nullptr;
    int64_t last_h2d = 0, last_d2h = 0;  // bytes the last agar_step_host call moved over PCIe
};

namespace {

size_t obs_elems(const AgarEnv* e) {
    const AgarConfig& c = e->p.c;
    return (size_t)c.num_arenas * c.num_agents * e->p.C * c.grid_size * c.grid_size;
}

// remembers on which caller stream work is in flight (see agar_step_host)
void note_stream(AgarEnv* e, cudaStream_t st) {
    if (e->pending && st != e->last_stream) e->pending_multi = true;
    e->last_stream = st;
    e->pending = true;
}

int launch(AgarEnv* e, KArgs& a, int n_records, cudaStream_t st, bool caller_stream = true) {
    a.n_records = n_records;
    const int N = e->p.c.num_arenas;
    bool ordered = e->order_enabled && e->d_order && (a.flags & F_STEP) && !(a.flags & (F_RESET | F_READONLY)) &&
                   n_records == N && a.arena0 == 0 && N >= 2048;
    if (ordered) {  // the order buffers rotate from launch to launch: a launch captured into a CUDA graph would replay with
        // one fixed triple (its output list never cleared), so captured launches keep the identity order
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { ordered = false; (void)cudaGetLastError(); }
    }
    if (ordered) {  // the order written by the last full step launch is consumed by this one, which writes the next buffer
        // and clears the header of the third (the one the launch after this writes; nothing else touches it meanwhile)
        const int in = e->order_cur, out = in < 0 ? 0 : (in + 1) % 3, clr = (out + 1) % 3;
        const size_t stride = ((size_t)N + 3) / 4 * 4;  // 16-byte aligned lists and headers
        auto buf = [&](int i) { return e->d_order + (size_t)i * (8 + 4 * stride); };
        if (in >= 0) a.order_in = buf(in);
        a.order_out = buf(out);
        a.order_clear = buf(clr);
        a.order_stride = (int32_t)stride;
        e->order_cur = out;
    }
    e->kernel<<<n_records, e->threads, e->smem, st>>>(e->p, a);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    if (caller_stream) note_stream(e, st);
    return 0;
}

int host_threads_default() {
    if (const char* v = std::getenv("AGAR_HOST_THREADS")) { const int n = std::atoi(v); if (n >= 1 && n <= kMaxWorkers) return n; }
    cpu_set_t set;
    int avail = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char* v = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::atoi(v) > 0 ? std::atoi(v) : 1;  // one process per GPU share the cores
    int n = avail / ranks;
    if (n >= 4) n -= 1;  // the calling thread polls the chunk events meanwhile: leave it a core (measured: 15 workers beat 16 on 16 cores)
    return n < 1 ? 1 : (n > 32 ? 32 : n);
}

KArgs base_args(AgarEnv* e) {
    KArgs a;
    std::memset(&a, 0, sizeof(a));
    a.state = e->d_state;
    a.src_state = e->d_state;
    a.events = e->d_events;
    a.ev_counts = e->d_ev_counts;
    a.hash = e->persist_hash ? e->d_hash : nullptr;  // (AGAR_NO_PERSIST: rebuild the pellet hash every step)
    return a;
}

}  // namespace

extern "C" {

const char* agar_last_error(void) { return g_err.c_str(); }
int agar_abi_version(void) { return AGAR_ABI_VERSION; }

int agar_config_default(AgarConfig* cfg) {
    if (!cfg) return fail("cfg is NULL");
    agar_spec_config_default(cfg);
    return 0;
}

int agar_layout_compute(const AgarConfig* cfg, AgarLayout* out) {
    if (!cfg || !out) return fail("NULL argument");
    if (const char* m = agar_spec_config_check(cfg)) return fail(m);
    agar_spec_layout(cfg, out);
    return 0;
}

int agar_obs_channels(const AgarConfig* cfg) { return cfg ? agar_spec_obs_channels(cfg) : 0; }

int agar_destroy(AgarEnv* e);

static int create_impl(const AgarConfig* cfg, int device, AgarEnv* e, const cudaDeviceProp& prop) {
    e->device = device;
    e->persist_hash = std::getenv("AGAR_NO_PERSIST") == nullptr;
    e->host_profile = std::getenv("AGAR_HOST_PROFILE") != nullptr;
    DevParams& p = e->p;
    std::memset(&p, 0, sizeof(p));
    p.c = *cfg;
    agar_spec_layout(cfg, &p.L);
    p.C = agar_spec_obs_channels(cfg);
    p.K = p.L.num_players;
    p.KC = p.L.cells_total;
    const int G = cfg->grid_size;
    const double box_d = (double)cfg->view_size / (double)G;  // features/extractors.py:19
    p.box = (float)box_d;
    p.hash_inv = (float)HASH_W / cfg->arena_size;
    p.hash_bs = cfg->arena_size / (float)HASH_W;
    {
        int ex = 0;
        p.box_inv = (std::frexp(p.box, &ex) == 0.5f) ? 1.0f / p.box : 0.0f;  // power of two: x * (1/box) == x / box
        p.raster_lim = (float)(G / 2 + 1) * p.box * 1.0001f;
    }
    for (int i = 0; i < G; ++i) p.oob_off[i] = (float)((double)(i - G / 2) * box_d);
    e->smem = smem_bytes(p.L, G);
    if (e->smem > (size_t)prop.sharedMemPerBlockOptin) {
        return fail("arena record + observation tile do not fit in shared memory (" + std::to_string(e->smem) + " B)");
    }
    // CTAs per SM that the shared memory allows (1 KB per CTA is reserved by the system)
    const int fit = (int)((size_t)prop.sharedMemPerMultiprocessor / (e->smem + 1024));
    e->min_blocks = fit >= 8 ? 8 : (fit == 7 ? 7 : (fit == 6 ? 6 : 5));
    if (const char* mb = std::getenv("AGAR_MIN_BLOCKS")) {  // tuning switch: force a register budget (5..8 CTAs/SM)
        const int v = std::atoi(mb);
        if (v >= 5 && v <= 8) e->min_blocks = v;
    }
    const bool generic_only = std::getenv("AGAR_GENERIC_KERNEL") != nullptr;  // tuning/test switch: never specialise
#ifdef AGAR_EXTRA_CV
    // a per-config copy of the library (build.build_specialised) carries only its own specialisation
    if (!generic_only && matches<CvExtra>(*cfg, p.L)) { e->kernel = agar_step_kernel<kThreads, AGAR_EXTRA_CV_BLOCKS, CvExtra>; e->kernel_name = "extra"; }
#else
    if (!generic_only && matches<CvCfg2>(*cfg, p.L)) {
        e->kernel = agar_step_kernel<kThreads, 7, CvCfg2>; e->kernel_name = "cfg2";
        if (const char* th = std::getenv("AGAR_THREADS")) {  // tuning switch: threads per arena
            if (std::atoi(th) == 64) { e->kernel = agar_step_kernel<64, 7, CvCfg2>; e->threads = 64; }
        }
    }
    else if (!generic_only && matches<CvCfg1>(*cfg, p.L)) {
        e->kernel_name = "cfg1";
        const char* mb = std::getenv("AGAR_MIN_BLOCKS");
        const int v = mb ? std::atoi(mb) : 10;  // measured on B200: 8 CTAs/SM 59.3 M, 10 (<=51 regs) 65.1 M, 12 (40 regs) 64.2 M env-steps/s
        e->kernel = v == 8 ? agar_step_kernel<kThreads, 8, CvCfg1> : (v == 12 ? agar_step_kernel<kThreads, 12, CvCfg1> : agar_step_kernel<kThreads, 10, CvCfg1>);
    }
    else if (!generic_only && matches<CvCfg4>(*cfg, p.L)) { e->kernel = agar_step_kernel<kThreads, 8, CvCfg4>; e->kernel_name = "cfg4"; }
    else if (!generic_only && matches<CvA2C>(*cfg, p.L)) { e->kernel = agar_step_kernel<kThreads, 6, CvA2C>; e->kernel_name = "a2c"; }
    else if (!generic_only && matches<CvA2CTest>(*cfg, p.L)) { e->kernel = agar_step_kernel<kThreads, 6, CvA2CTest>; e->kernel_name = "a2c_test"; }
    else if (!generic_only && matches<CvDQN>(*cfg, p.L)) { e->kernel = agar_step_kernel<kThreads, 8, CvDQN>; e->kernel_name = "dqn"; }
#endif
    else {
        e->kernel_name = "generic";
        switch (e->min_blocks) {
            case 8: e->kernel = agar_step_kernel<kThreads, 8, RtCfg>; break;
            case 7: e->kernel = agar_step_kernel<kThreads, 7, RtCfg>; break;
            case 6: e->kernel = agar_step_kernel<kThreads, 6, RtCfg>; break;
            default: e->kernel = agar_step_kernel<kThreads, 5, RtCfg>; break;
        }
    }
    CUDA_OK(cudaFuncSetAttribute(e->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    const size_t state_bytes = (size_t)cfg->num_arenas * p.L.words * sizeof(float);
    CUDA_OK(cudaMalloc(&e->d_state, state_bytes));
    CUDA_OK(cudaMemsetAsync(e->d_state, 0, state_bytes, 0));
    p.hash_words = (int)hash_record_words(p.L);
    p.hash_epoch = 1;
    CUDA_OK(cudaMalloc(&e->d_hash, (size_t)cfg->num_arenas * p.hash_words * sizeof(int32_t)));
    CUDA_OK(cudaMemsetAsync(e->d_hash, 0, (size_t)cfg->num_arenas * p.hash_words * sizeof(int32_t), 0));  // epoch 0 = stale
    if (cfg->event_capacity > 0) {
        CUDA_OK(cudaMalloc(&e->d_events, (size_t)cfg->num_arenas * cfg->event_capacity * 4 * sizeof(int32_t)));
        CUDA_OK(cudaMalloc(&e->d_ev_counts, (size_t)cfg->num_arenas * sizeof(int32_t)));
        CUDA_OK(cudaMemsetAsync(e->d_ev_counts, 0, (size_t)cfg->num_arenas * sizeof(int32_t), 0));
    }
    e->order_enabled = std::getenv("AGAR_NO_ORDER") == nullptr;
    const size_t order_words = 3 * (8 + 4 * (((size_t)cfg->num_arenas + 3) / 4 * 4));
    CUDA_OK(cudaMalloc(&e->d_order, order_words * sizeof(int32_t)));
    CUDA_OK(cudaMemsetAsync(e->d_order, 0, order_words * sizeof(int32_t), 0));
    CUDA_OK(cudaEventCreateWithFlags(&e->order_event, cudaEventDisableTiming));
    // the clears ran on the legacy stream, which non-blocking caller streams do not wait for: finish them here
    CUDA_OK(cudaStreamSynchronize(0));
    return 0;
}

int agar_create(const AgarConfig* cfg, int device, AgarEnv** out) {
    if (!cfg || !out) return fail("NULL argument");
    if (const char* m = agar_spec_config_check(cfg)) return fail(std::string("bad config: ") + m);
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("no such CUDA device");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("libagar_b200 is built for sm_100a (B200) only");
    AgarEnv* e = new (std::nothrow) AgarEnv();
    if (!e) return fail("out of host memory");
    if (int rc = create_impl(cfg, device, e, prop)) {  // g_err is set; release whatever was allocated
        agar_destroy(e);
        return rc;
    }
    *out = e;
    return 0;
}


int agar_destroy(AgarEnv* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    if (e->host_profile && e->prof_calls > 0)
        std::fprintf(stderr, "[agar_step_host] %lld calls, us per call: enqueue %.0f | wait for lists %.0f | host threads %.0f | final syncs %.0f\n",
                     (long long)e->prof_calls, e->prof_us[0] / e->prof_calls, e->prof_us[1] / e->prof_calls, e->prof_us[2] / e->prof_calls,
                     e->prof_us[3] / e->prof_calls);
    if (e->host_profile && e->prof_calls > 0 && e->pool)
        std::fprintf(stderr, "[agar_step_host] kilo-cycles from the last chunk's publication to the first / median / last worker's finish: %.0f / %.0f / %.0f\n",
                     e->prof_spread[0] / e->prof_calls, e->prof_spread[1] / e->prof_calls, e->prof_spread[2] / e->prof_calls);
    if (e->host_profile && e->prof_calls > 0 && e->pool)
        std::fprintf(stderr, "[agar_step_host] worker kilo-cycles per call, summed over the workers: waiting %lld | plan %lld | prefetch %lld | in-place %lld | full rewrite %lld\n",
                     e->prof_tsc[0].load() / e->prof_calls / 1000, e->prof_tsc[1].load() / e->prof_calls / 1000, e->prof_tsc[2].load() / e->prof_calls / 1000,
                     e->prof_tsc[3].load() / e->prof_calls / 1000, e->prof_tsc[4].load() / e->prof_calls / 1000);
    cudaFree(e->d_order); cudaFree(e->d_state); cudaFree(e->d_events); cudaFree(e->d_ev_counts); cudaFree(e->d_hash);
    cudaFree(e->d_tab_xy); cudaFree(e->d_tab_act);
    cudaFree(e->d_target); cudaFree(e->d_act); cudaFree(e->d_obs); cudaFree(e->d_reward); cudaFree(e->d_done);
    if (e->host_stream) cudaStreamDestroy(e->host_stream);
    if (e->host_stream2) cudaStreamDestroy(e->host_stream2);
    if (e->host_event) cudaEventDestroy(e->host_event);
    if (e->order_event) cudaEventDestroy(e->order_event);
    delete e->pool;
    delete[] e->take;
    cudaFree(e->d_entries); cudaFree(e->d_meta);
    cudaFreeHost(e->h_entries); cudaFreeHost(e->h_meta);
    for (int k = 0; k < kHostChunks; ++k)
        if (e->chunk_done[k]) cudaEventDestroy(e->chunk_done[k]);
    delete e;
    return 0;
}

int agar_get_config(const AgarEnv* e, AgarConfig* out) {
    if (!e || !out) return fail("NULL argument");
    *out = e->p.c;
    return 0;
}

int agar_get_layout(const AgarEnv* e, AgarLayout* out) {
    if (!e || !out) return fail("NULL argument");
    *out = e->p.L;
    return 0;
}

int agar_seed(AgarEnv* e, uint64_t seed) {
    if (!e) return fail("env is NULL");
    CUDA_OK(cudaSetDevice(e->device));
    // no stream argument: the call is synchronous with respect to everything in flight on the device (steps the
    // caller enqueued on non-blocking streams included), before and after the tick counters are cleared
    CUDA_OK(cudaDeviceSynchronize());
    e->p.seed_lo = (uint32_t)seed;
    e->p.seed_hi = (uint32_t)(seed >> 32);
    e->p.hash_epoch += 1;  // conservative: persisted hashes are rebuilt
    // tick counters restart from 0: one strided 4-byte memset per arena record
    CUDA_OK(cudaMemset2DAsync(e->d_state + e->p.L.tick, (size_t)e->p.L.words * sizeof(float), 0, sizeof(int32_t),
                              (size_t)e->p.c.num_arenas, 0));
    CUDA_OK(cudaStreamSynchronize(0));
    return 0;
}

int agar_reset(AgarEnv* e, const uint8_t* d_reset_mask, float* d_obs, void* stream) {
    if (!e) return fail("env is NULL");
    CUDA_OK(cudaSetDevice(e->device));
    KArgs a = base_args(e);  // the kernel marks the persisted pellet hash of every arena it resets as stale
    a.flags = F_RESET | (d_obs ? F_OBS : 0);
    a.reset_mask = d_reset_mask;
    a.obs = d_obs;
    return launch(e, a, e->p.c.num_arenas, (cudaStream_t)stream);
}

int agar_step(AgarEnv* e, const float* d_target_xy, const int32_t* d_act, float* d_obs, float* d_reward,
              uint8_t* d_done, void* stream) {
    if (!e) return fail("env is NULL");
    if (!d_target_xy || !d_act || !d_reward || !d_done) return fail("agar_step: NULL action/reward/done buffer");
    CUDA_OK(cudaSetDevice(e->device));
    KArgs a = base_args(e);
    a.flags = F_STEP | (d_obs ? F_OBS : 0);
    a.target_xy = d_target_xy; a.act = d_act; a.obs = d_obs; a.reward = d_reward; a.done = d_done;
    return launch(e, a, e->p.c.num_arenas, (cudaStream_t)stream);
}

int agar_set_action_table(AgarEnv* e, const float* h_xy, const int32_t* h_act, int32_t n) {
    if (!e || !h_xy || !h_act) return fail("NULL argument");
    if (n < 1 || n > 65536) return fail("n_actions out of range");
    CUDA_OK(cudaSetDevice(e->device));
    cudaFree(e->d_tab_xy); cudaFree(e->d_tab_act);
    e->d_tab_xy = nullptr; e->d_tab_act = nullptr;
    CUDA_OK(cudaMalloc(&e->d_tab_xy, (size_t)n * 2 * sizeof(float)));
    CUDA_OK(cudaMalloc(&e->d_tab_act, (size_t)n * sizeof(int32_t)));
    CUDA_OK(cudaMemcpy(e->d_tab_xy, h_xy, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->d_tab_act, h_act, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    e->n_actions = n;
    return 0;
}

int agar_step_indexed(AgarEnv* e, const int32_t* d_action_index, float* d_obs, float* d_reward, uint8_t* d_done,
                      void* stream) {
    if (!e) return fail("env is NULL");
    if (!e->d_tab_xy) return fail("agar_step_indexed: call agar_set_action_table first");
    if (!d_action_index || !d_reward || !d_done) return fail("agar_step_indexed: NULL buffer");
    CUDA_OK(cudaSetDevice(e->device));
    KArgs a = base_args(e);
    a.flags = F_STEP | (d_obs ? F_OBS : 0);
    a.act_index = d_action_index; a.tab_xy = e->d_tab_xy; a.tab_act = e->d_tab_act; a.n_actions = e->n_actions;
    a.obs = d_obs; a.reward = d_reward; a.done = d_done;
    return launch(e, a, e->p.c.num_arenas, (cudaStream_t)stream);
}

int agar_observe(AgarEnv* e, float* d_obs, void* stream) {
    if (!e || !d_obs) return fail("NULL argument");
    CUDA_OK(cudaSetDevice(e->device));
    KArgs a = base_args(e);
    a.flags = F_OBS | F_READONLY;
    a.obs = d_obs;
    return launch(e, a, e->p.c.num_arenas, (cudaStream_t)stream);
}

int agar_observe_records(AgarEnv* e, const float* d_state, int32_t n_records, float* d_obs, void* stream) {
    if (!e || !d_obs || !d_state) return fail("NULL argument");
    if (n_records < 1) return fail("n_records must be >= 1");
    if (((uintptr_t)d_state & 15) != 0) return fail("d_state must be 16-byte aligned");
    CUDA_OK(cudaSetDevice(e->device));
    KArgs a = base_args(e);
    a.flags = F_OBS | F_READONLY;
    a.src_state = d_state;
    a.obs = d_obs;
    return launch(e, a, n_records, (cudaStream_t)stream);
}

int agar_get_state(AgarEnv* e, float* d_dst, void* stream) {
    if (!e || !d_dst) return fail("NULL argument");
    CUDA_OK(cudaSetDevice(e->device));
    CUDA_OK(cudaMemcpyAsync(d_dst, e->d_state, (size_t)e->p.c.num_arenas * e->p.L.words * sizeof(float),
                            cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    note_stream(e, (cudaStream_t)stream);
    return 0;
}

int agar_set_state(AgarEnv* e, const float* d_src, void* stream) {
    if (!e || !d_src) return fail("NULL argument");
    CUDA_OK(cudaSetDevice(e->device));
    CUDA_OK(cudaMemcpyAsync(e->d_state, d_src, (size_t)e->p.c.num_arenas * e->p.L.words * sizeof(float),
                            cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    e->p.hash_epoch += 1;  // the pellets may have changed: every persisted hash is stale
    note_stream(e, (cudaStream_t)stream);
    return 0;
}

int agar_events(AgarEnv* e, int32_t* d_events, int32_t* d_counts, void* stream) {
    if (!e || !d_events || !d_counts) return fail("NULL argument");
    const int cap = e->p.c.event_capacity;
    if (cap <= 0) return fail("event log disabled (event_capacity == 0)");
    CUDA_OK(cudaSetDevice(e->device));
    const size_t N = (size_t)e->p.c.num_arenas;
    CUDA_OK(cudaMemcpyAsync(d_events, e->d_events, N * cap * 4 * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    CUDA_OK(cudaMemcpyAsync(d_counts, e->d_ev_counts, N * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    note_stream(e, (cudaStream_t)stream);
    return 0;
}

int agar_step_host(AgarEnv* e, const float* h_target_xy, const int32_t* h_act, float* h_obs, float* h_reward,
                   uint8_t* h_done) {
    if (!e) return fail("env is NULL");
    if (!h_target_xy || !h_act || !h_reward || !h_done) return fail("agar_step_host: NULL buffer");
    CUDA_OK(cudaSetDevice(e->device));
    const size_t N = (size_t)e->p.c.num_arenas, A = (size_t)e->p.c.num_agents, NA = N * A;
    if (!e->host_stream) {  // first call: private streams and staging buffers; the handle keeps them only if all succeed
        cudaStream_t s1 = nullptr, s2 = nullptr;
        cudaEvent_t ev = nullptr;
        float* t = nullptr; int32_t* ac = nullptr; float* rw = nullptr; uint8_t* dn = nullptr;
        cudaError_t err = cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
        if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaMalloc(&t, NA * 2 * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&ac, NA * sizeof(int32_t));
        if (err == cudaSuccess) err = cudaMalloc(&rw, NA * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&dn, NA);
        if (err != cudaSuccess) {
            if (s1) cudaStreamDestroy(s1);
            if (s2) cudaStreamDestroy(s2);
            if (ev) cudaEventDestroy(ev);
            cudaFree(t); cudaFree(ac); cudaFree(rw); cudaFree(dn);
            return fail(std::string("agar_step_host: staging allocation failed: ") + cudaGetErrorString(err));
        }
        e->host_stream = s1; e->host_stream2 = s2; e->host_event = ev;
        e->d_target = t; e->d_act = ac; e->d_reward = rw; e->d_done = dn;
    }
    if (h_obs && !e->d_obs) CUDA_OK(cudaMalloc(&e->d_obs, obs_elems(e) * sizeof(float)));
    const size_t per_agent = (size_t)e->p.C * e->p.c.grid_size * e->p.c.grid_size, per_arena = A * per_agent;
    // Compact copy (default for batches of at least 256 agents; AGAR_HOST_DENSE=1 keeps the plain image copy): the
    // images stay on the device, the host receives each agent's centroid and occupied bins and rebuilds the dense
    // grids with its own threads.  201 MB -> 8.5 MB over PCIe per step of 4096 cfg2 arenas.
    if (h_obs && !e->entry_cap && NA >= 256 && !std::getenv("AGAR_HOST_DENSE")) {
        int want = kEntryCap;
        if (const char* v = std::getenv("AGAR_ENTRY_CAP")) { const int n = std::atoi(v); if (n >= 16 && n <= 4096) want = n; }  // tuning switch
        const int cap = per_agent < (size_t)want ? (int)per_agent : want;
        cudaError_t err = cudaMalloc(&e->d_entries, NA * cap * sizeof(ObsEntry));
        if (err == cudaSuccess) err = cudaMalloc(&e->d_meta, NA * 4 * sizeof(float));
        if (err == cudaSuccess) err = cudaMallocHost(&e->h_entries, 2 * NA * cap * sizeof(ObsEntry));
        if (err == cudaSuccess) err = cudaMallocHost(&e->h_meta, 2 * NA * 4 * sizeof(float));
        for (int k = 0; k < kHostChunks && err == cudaSuccess; ++k) err = cudaEventCreateWithFlags(&e->chunk_done[k], cudaEventDisableTiming);
        if (err != cudaSuccess) return fail(std::string("agar_step_host: compact-copy allocation failed: ") + cudaGetErrorString(err));
        e->pool = new (std::nothrow) HostPool(host_threads_default());
        e->take = new (std::nothrow) TakeCounter[kHostChunks * kMaxWorkers];
        if (!e->pool || !e->take) return fail("out of host memory");
        e->entry_cap = cap;
        e->host_compact = true;
    }
    const bool compact = h_obs && e->host_compact;
    CallerPin caller_pin(compact && e->pool ? e->pool->caller_cpu() : -1);
    // this call fills the half of the pinned list buffers the previous one did not; with a persistent h_obs the previous
    // half says what the buffer holds and the grids are updated in place
    const int cur = e->hbuf ^ 1;
    const bool delta = compact && e->host_persistent && e->delivered == h_obs;
    e->delivered = nullptr;  // (set again once this call has completed)
    const size_t half = compact ? (size_t)cur * NA : 0, other = compact ? (size_t)(cur ^ 1) * NA : 0;
    ObsEntry* const h_ent = e->h_entries + half * (size_t)e->entry_cap;
    float* const h_met = e->h_meta + half * 4;
    const ObsEntry* const p_ent = e->h_entries + other * (size_t)e->entry_cap;
    const float* const p_met = e->h_meta + other * 4;
    // an agent's record in the meta arrays: {centroid x, centroid y, has cells, length of its list as int32 bits}
    const auto h_cnt = [h_met](size_t g) { return reinterpret_cast<const int32_t*>(h_met)[g * 4 + 3]; };
    const auto p_cnt = [p_met](size_t g) { return reinterpret_cast<const int32_t*>(p_met)[g * 4 + 3]; };
    // The private streams are non-blocking: order them after whatever the caller enqueued on its own stream(s)
    // (agar_reset, agar_step, agar_set_state ...) since the previous host-buffer step.
    if (e->pending) {
        if (e->pending_multi) CUDA_OK(cudaDeviceSynchronize());
        else {
            CUDA_OK(cudaEventRecord(e->order_event, e->last_stream));
            CUDA_OK(cudaStreamWaitEvent(e->host_stream, e->order_event, 0));
        }
        e->pending = false; e->pending_multi = false;
    }
    cudaStream_t st = e->host_stream, st2 = e->host_stream2;
    const bool prof = e->host_profile;
    auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = prof ? now_us() : 0.0;
    auto lap = [&](int slot) { if (prof) { const double t = now_us(); e->prof_us[slot] += t - t_mark; t_mark = t; } };
    e->last_h2d = (int64_t)(NA * 2 * sizeof(float) + NA * sizeof(int32_t));
    e->last_d2h = (int64_t)(NA * sizeof(float) + NA);
    CUDA_OK(cudaMemcpyAsync(e->d_target, h_target_xy, NA * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(e->d_act, h_act, NA * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // The batch is stepped in chunks so that the device-to-host copy of one chunk (copy stream), and the host's work on
    // it, overlap the kernel of the next (compute stream).
    const int chunks = (h_obs && N >= 1024) ? kHostChunks : 1;
    const int cap = e->entry_cap;
    // One job per call: the workers are woken now (their wake-up overlaps the first chunk's kernel) and go through the
    // chunks in order, each as soon as `ready` says its lists have landed; kAbort releases them when the call fails.
    constexpr int kAbort = 1 << 30;
    std::atomic<int> ready{0};
    std::function<void(int, int)> job;
    struct JobGuard {
        HostPool* pool; std::atomic<int>* ready; bool active;
        ~JobGuard() { if (active) { ready->store(kAbort, std::memory_order_release); pool->wait(); } }
    } guard{e->pool, &ready, false};
    if (compact) {
        const GridGeo geo{e->p.c.grid_size, e->p.C, e->p.K, e->p.c.obs_flags, e->p.c.arena_size, e->p.oob_off};  // (oob_off lives in the handle)
        const int sentmask = sentinel_mask(geo);
        std::atomic<int>* rdy = &ready;
        std::atomic<long long>* ptsc = prof ? e->prof_tsc : nullptr;
        long long* pfin = e->prof_fin;
        TakeCounter* take = e->take;
        for (int i = 0; i < kHostChunks * kMaxWorkers; ++i) take[i].v.store(0, std::memory_order_relaxed);
        job = [=](int w, int nw) {
            DeltaPlan plan[2];
            long long tsc[5] = {0, 0, 0, 0, 0}, t_last = ptsc ? (long long)__rdtsc() : 0;
            auto tick = [&](int slot) { if (ptsc) { const long long t = (long long)__rdtsc(); tsc[slot] += t - t_last; t_last = t; } };
            auto in_place = [&](size_t g) { return delta && h_cnt(g) <= cap && p_cnt(g) <= cap; };  // (an image copied whole has no list)
            for (int k = 0; k < chunks; ++k) {
                const size_t g0 = (size_t)(N * k / chunks) * A, g1 = (size_t)(N * (k + 1) / chunks) * A;
                for (int spins = 0;; ++spins) {
                    const int r = rdy->load(std::memory_order_acquire);
                    if (r >= kAbort) return;
                    if (r > k) break;
                    if ((spins & 1023) == 1023) std::this_thread::yield(); else _mm_pause();
                }
                tick(0);
                // Worker v's list of this chunk is the agents g0 + v, g0 + v + nw, ...; `take[k][v]` counts how many of
                // them have been taken.  A worker goes through its own list (the same agents every step: their lines are
                // in its core's L2), then takes from the lists of the others, so that a worker whose core is shared
                // with something else for a while does not hold the call back (measured: the last worker finished
                // 2-4x later than the median one).
                constexpr size_t kEnd = ~(size_t)0;
                // (taken kTake at a time: the locked add waits for the stores of the agent just written to drain, which
                // costs ~1 k cycles when paid per agent)
                constexpr int kTake = 8;
                int victim = w;
                size_t gb = 0, ge = 0;  // the agents in hand: gb, gb + nw, ... below ge
                auto fetch = [&]() -> size_t {
                    if (gb < ge) { const size_t g = gb; gb += (size_t)nw; return g; }
                    for (int tries = 0; tries < nw; ++tries) {
                        const size_t i0 = (size_t)take[k * kMaxWorkers + victim].v.fetch_add(kTake, std::memory_order_relaxed);
                        const size_t g = g0 + (size_t)victim + (size_t)nw * i0;
                        if (g < g1) {
                            const size_t lim = g + (size_t)nw * kTake;
                            ge = lim < g1 ? lim : g1;
                            gb = g + (size_t)nw;
                            return g;
                        }
                        victim = victim + 1 == nw ? 0 : victim + 1;
                    }
                    return kEnd;
                };
                int pi = 0;
                size_t g = fetch();
                if (g != kEnd && in_place(g)) delta_plan(geo, p_met + g * 4, h_met + g * 4, sentmask, plan[0]);
                while (g != kEnd) {
                    const size_t gn = fetch();  // plan the agent after this one and start its misses now
                    if (gn != kEnd && in_place(gn)) {
                        delta_plan(geo, p_met + gn * 4, h_met + gn * 4, sentmask, plan[pi ^ 1]);
                        tick(1);
                        delta_prefetch(geo, h_obs + gn * per_agent, plan[pi ^ 1], p_ent + gn * cap, p_cnt(gn), h_ent + gn * cap, h_cnt(gn));
                        tick(2);
                    }
                    const int cnt = h_cnt(g);
                    if (cnt <= cap) {  // (an overflowed list: the calling thread has the image copied whole)
                        if (in_place(g)) { delta_apply(geo, h_obs + g * per_agent, plan[pi], p_ent + g * cap, p_cnt(g), h_ent + g * cap, cnt, h_met + g * 4); tick(3); }
                        else { expand_agent(geo, h_obs + g * per_agent, h_ent + g * cap, cnt, h_met + g * 4); tick(4); }
                    }
                    g = gn;
                    pi ^= 1;
                }
            }
            if (ptsc) {
                for (int i = 0; i < 5; ++i) ptsc[i].fetch_add(tsc[i], std::memory_order_relaxed);
                if (w < 64) pfin[w] = (long long)__rdtsc();
            }
        };
        e->pool->start(job);
        guard.active = true;
    }
    for (int k = 0; k < chunks; ++k) {
        const int a0 = (int)(N * k / chunks), a1 = (int)(N * (k + 1) / chunks);
        KArgs a = base_args(e);
        a.flags = F_STEP | (h_obs ? F_OBS : 0);
        a.target_xy = e->d_target; a.act = e->d_act; a.obs = h_obs ? e->d_obs : nullptr;
        a.obs_meta = compact ? e->d_meta : nullptr;
        a.reward = e->d_reward; a.done = e->d_done;
        a.arena0 = a0;
        if (int rc = launch(e, a, a1 - a0, st, false)) return rc;
        if (!h_obs) continue;
        const size_t g0 = (size_t)a0 * A, gn = (size_t)(a1 - a0) * A;  // agents of this chunk
        if (compact) {  // (on the compute stream: beside the next chunk's step kernel, on the copy stream, the lists arrive later - measured)
            obs_compact_kernel<<<(unsigned)gn, 256, 0, st>>>(e->d_obs + g0 * per_agent, (int)per_agent, cap, e->d_entries + g0 * cap,
                                                             reinterpret_cast<int32_t*>(e->d_meta + g0 * 4) + 3, 4);
            CUDA_OK(cudaGetLastError());
            e->launches += 1;
        }
        CUDA_OK(cudaEventRecord(e->host_event, st));
        CUDA_OK(cudaStreamWaitEvent(st2, e->host_event, 0));
        if (compact) {
            CUDA_OK(cudaMemcpyAsync(h_ent + g0 * cap, e->d_entries + g0 * cap, gn * cap * sizeof(ObsEntry), cudaMemcpyDeviceToHost, st2));
            CUDA_OK(cudaMemcpyAsync(h_met + g0 * 4, e->d_meta + g0 * 4, gn * 4 * sizeof(float), cudaMemcpyDeviceToHost, st2));
            CUDA_OK(cudaEventRecord(e->chunk_done[k], st2));
            e->last_d2h += (int64_t)(gn * cap * sizeof(ObsEntry) + gn * 4 * sizeof(float));
        } else {
            e->last_d2h += (int64_t)((size_t)(a1 - a0) * per_arena * sizeof(float));
            CUDA_OK(cudaMemcpyAsync(h_obs + a0 * per_arena, e->d_obs + a0 * per_arena, (size_t)(a1 - a0) * per_arena * sizeof(float),
                                    cudaMemcpyDeviceToHost, st2));
        }
    }
    CUDA_OK(cudaMemcpyAsync(h_reward, e->d_reward, NA * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(h_done, e->d_done, NA, cudaMemcpyDeviceToHost, st));
    lap(0);
    if (compact) {
        for (int k = 0; k < chunks; ++k) {  // hand each chunk to the workers as soon as its lists have landed
            const size_t g0 = (size_t)(N * k / chunks) * A, g1 = (size_t)(N * (k + 1) / chunks) * A;
            CUDA_OK(cudaEventSynchronize(e->chunk_done[k]));
            ready.store(k + 1, std::memory_order_release);
            for (size_t g = g0; g < g1; ++g)  // an overflowed list: the image is copied whole (the workers skip the agent)
                if (h_cnt(g) > cap) {
                    CUDA_OK(cudaMemcpyAsync(h_obs + g * per_agent, e->d_obs + g * per_agent, per_agent * sizeof(float), cudaMemcpyDeviceToHost, st2));
                    e->last_d2h += (int64_t)(per_agent * sizeof(float));
                }
        }
        lap(1);
        const long long t_pub = prof ? (long long)__rdtsc() : 0;
        e->pool->wait();
        guard.active = false;
        lap(2);
        if (prof) {
            std::vector<long long> f(e->prof_fin, e->prof_fin + std::min(64, e->pool->size()));
            std::sort(f.begin(), f.end());
            e->prof_spread[0] += (double)(f.front() - t_pub) / 1e3; e->prof_spread[1] += (double)(f[f.size() / 2] - t_pub) / 1e3;
            e->prof_spread[2] += (double)(f.back() - t_pub) / 1e3;
        }
    }
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaStreamSynchronize(st2));
    lap(3);
    if (prof) e->prof_calls += 1;
    if (compact) { e->hbuf = cur; e->delivered = h_obs; }
    return 0;
}

int agar_host_obs_persistent(AgarEnv* e, int on) {
    if (!e) return fail("env is NULL");
    e->host_persistent = on != 0;
    e->delivered = nullptr;  // the next call rewrites every grid in full
    return 0;
}

int agar_gae(const float* d_reward, const float* d_value, const uint8_t* d_done, const float* d_end_value,
             double gamma, double lam, float* d_adv, float* d_ret, int32_t T, int32_t N, void* stream) {
    if (T < 0 || N < 0) return fail("agar_gae: negative size");
    if (T == 0 || N == 0) return 0;  // test/returns_test.py:15-19: empty rollouts are legal
    if (!d_reward) return fail("agar_gae: d_reward is NULL");
    if (d_adv && !d_value) return fail("agar_gae: advantages need d_value");
    if (!d_adv && !d_ret) return fail("agar_gae: nothing to compute");
    const float g = (float)gamma, cgl = (float)(gamma * lam);
    // tuning switches (defaults are the measured best on B200, profiles/): steps per register batch and block size.
    // Variants with 2 or 4 columns per thread (64/128-bit loads) were measured slower: too few threads (55 / 60 us vs 38 us).
    static const int tune_u = std::getenv("AGAR_GAE_U") ? std::atoi(std::getenv("AGAR_GAE_U")) : 16;
    static const int tune_b = std::getenv("AGAR_GAE_BLOCK") ? std::atoi(std::getenv("AGAR_GAE_BLOCK")) : 32;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = (tune_b == 64 || tune_b == 128) ? tune_b : 32, blocks = (N + threads - 1) / threads;
    static const int tune_nb = std::getenv("AGAR_GAE_NB") ? std::atoi(std::getenv("AGAR_GAE_NB")) : 2;
    // Default: the TMA-fed kernel (producer warp + ring of 5 shared-memory stages of 8 rows; measured 30.8 us = 72 % of
    // the HBM peak at [128, 65,536] including the launch, against 38-47 us for the register-ring kernel below; 3 stages
    // 31.8 us, 4 stages 31.3 us, 6 stages of 64 columns 31.6 us; profiles/r02_notes.txt).  It needs N % 16 == 0 and
    // 16-byte aligned rows.
    static const bool no_tma = std::getenv("AGAR_GAE_NO_TMA") != nullptr;
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (!no_tma && (N % 16) == 0 && al16(d_reward) && al16(d_value) && al16(d_done)) {
        static const char* tune = std::getenv("AGAR_GAE_TMA");  // tuning switch: "U,STAGES,W"
        int tu = 8, ts = 5, tw = 128;
        if (tune) std::sscanf(tune, "%d,%d,%d", &tu, &ts, &tw);
        // strip width: a multiple of 16 columns (bulk copies of the done bytes), at most the CTA's thread count, chosen so
        // that the strips fill per_sm resident CTAs on every SM evenly
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
        static const int per_sm = std::getenv("AGAR_GAE_PER_SM") ? std::atoi(std::getenv("AGAR_GAE_PER_SM")) : 4;
        auto strip_for = [&](int wmax) {
            int s0 = (N + sms * per_sm - 1) / (sms * per_sm);
            s0 = (s0 + 15) / 16 * 16;
            if (s0 < 32) s0 = 32;
            return s0 > wmax ? wmax : s0;
        };
#define GAE_TMA(U_, S_, W_) do { const int strip = strip_for(W_); agar_gae_tma_kernel<U_, S_, W_><<<(N + strip - 1) / strip, W_ + 32, 0, st>>>(d_reward, d_value, d_done, d_end_value, g, cgl, d_adv, d_ret, T, N, strip); } while (0)
        if (tu == 8 && ts == 4 && tw == 128) GAE_TMA(8, 4, 128);
        else if (tu == 10 && ts == 3 && tw == 128) GAE_TMA(10, 3, 128);
        else if (tu == 4 && ts == 6 && tw == 128) GAE_TMA(4, 6, 128);
        else if (tu == 8 && ts == 3 && tw == 64) GAE_TMA(8, 3, 64);
        else if (tu == 8 && ts == 6 && tw == 64) GAE_TMA(8, 6, 64);
        else if (tu == 8 && ts == 2 && tw == 256) GAE_TMA(8, 2, 256);
        else if (tu == 8 && ts == 3 && tw == 32) GAE_TMA(8, 3, 32);
        else if (tu == 8 && ts == 6 && tw == 32) GAE_TMA(8, 6, 32);
        else if (tu == 8 && ts == 3 && tw == 128) GAE_TMA(8, 3, 128);
        else GAE_TMA(8, 5, 128);
#undef GAE_TMA
        CUDA_OK(cudaGetLastError());
        return 0;
    }
#define GAE_LAUNCH(U_, NB_) agar_gae_kernel<U_, NB_><<<blocks, threads, 0, st>>>(d_reward, d_value, d_done, d_end_value, g, cgl, d_adv, d_ret, T, N)
    if (tune_u == 8 && tune_nb == 2) GAE_LAUNCH(8, 2);
    else if (tune_u == 8 && tune_nb == 3) GAE_LAUNCH(8, 3);
    else if (tune_u == 8 && tune_nb == 4) GAE_LAUNCH(8, 4);
    else if (tune_u == 8 && tune_nb == 6) GAE_LAUNCH(8, 6);
    else if (tune_u == 16 && tune_nb == 3) GAE_LAUNCH(16, 3);
    else if (tune_u == 16 && tune_nb == 4) GAE_LAUNCH(16, 4);
    else GAE_LAUNCH(16, 2);
#undef GAE_LAUNCH
    CUDA_OK(cudaGetLastError());
    return 0;
}

int agar_ram_feature_size(int32_t n_pellet, int32_t n_virus, int32_t n_food, int32_t n_other, int32_t n_cell) {
    if (n_pellet < 0 || n_virus < 0 || n_food < 0 || n_other < 0 || n_cell < 0) return -1;
    return 2 * (n_pellet + n_virus + n_food) + 5 * (1 + n_other) * n_cell;
}

int agar_observe_ram(AgarEnv* e, int32_t n_pellet, int32_t n_virus, int32_t n_food, int32_t n_other, int32_t n_cell,
                     float* d_out, void* stream) {
    if (!e || !d_out) return fail("NULL argument");
    const int size = agar_ram_feature_size(n_pellet, n_virus, n_food, n_other, n_cell);
    if (size <= 0) return fail("agar_observe_ram: bad feature counts");
    if (n_cell > AGAR_MAX_CELLS) return fail("agar_observe_ram: n_cell exceeds the 16 cell slots of a player");
    CUDA_OK(cudaSetDevice(e->device));
    RamParams rp;
    rp.n_pellet = n_pellet; rp.n_virus = n_virus; rp.n_food = n_food; rp.n_other = n_other; rp.n_cell = n_cell;
    rp.size = size; rp.filler = -1000.0f;
    const AgarConfig& c = e->p.c;
    int most = c.num_pellets > c.num_viruses ? c.num_pellets : c.num_viruses;
    if (c.max_foods > most) most = c.max_foods;
    rp.sort_cap = 1;
    while (rp.sort_cap < most) rp.sort_cap <<= 1;
    const size_t smem = (size_t)(rp.sort_cap + 64 + RAM_CAND_CAP) * sizeof(unsigned long long);
    CUDA_OK(cudaFuncSetAttribute(ram_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(ram_features_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ram_features_kernel<<<c.num_arenas * c.num_agents, 128, smem, (cudaStream_t)stream>>>(e->p, rp, e->d_state, d_out);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    note_stream(e, (cudaStream_t)stream);
    return 0;
}

int agar_screen_gray(const uint8_t* d_rgb, int64_t n_pixels, void* d_out, int32_t out_f64, void* stream) {
    if (n_pixels < 0) return fail("agar_screen_gray: negative size");
    if (n_pixels == 0) return 0;
    if (!d_rgb || !d_out) return fail("NULL argument");
    if (((uintptr_t)d_rgb & 15) != 0 || ((uintptr_t)d_out & 15) != 0) return fail("agar_screen_gray: buffers must be 16-byte aligned");
    const unsigned blocks = (unsigned)((n_pixels + GRAY_TILE - 1) / GRAY_TILE);  // a CTA per tile of 4096 pixels
    if (out_f64) screen_gray_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_rgb, (long long)n_pixels, (double*)d_out);
    else screen_gray_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_rgb, (long long)n_pixels, (float*)d_out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t agar_launch_count(const AgarEnv* e) { return e ? e->launches : 0; }

int agar_host_copy_bytes(const AgarEnv* e, int64_t* h2d, int64_t* d2h) {
    if (!e || !h2d || !d2h) return fail("NULL argument");
    *h2d = e->last_h2d; *d2h = e->last_d2h;
    return 0;
}
const char* agar_kernel_variant(const AgarEnv* e) { return e ? e->kernel_name : ""; }

int64_t agar_step_smem_bytes(const AgarConfig* cfg) {
    if (!cfg || agar_spec_config_check(cfg)) return -1;
    AgarLayout L;
    agar_spec_layout(cfg, &L);
    return (int64_t)smem_bytes(L, cfg->grid_size);
}

#ifdef AGAR_PHASE_CLOCKS
// tuning builds only (not declared in the header): cycles charged to each phase id since the last call
int agar_debug_phase_clocks(unsigned long long* h_out32) {
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpyFromSymbol(h_out32, g_phase_clocks, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    CUDA_OK(cudaMemcpyToSymbol(g_phase_clocks, zero, sizeof(zero)));
    return 0;
}
int agar_debug_arena_cycles(unsigned int* h_out8192) {
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpyFromSymbol(h_out8192, g_arena_cycles, sizeof(unsigned int) * 8192));
    return 0;
}
#endif

}  // extern "C"
