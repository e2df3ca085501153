// Multi-GPU film reduction over NVLink peer memory (msk_gpu_reduce_film).
//
// The path shards by sample range: every GPU renders its samples of EVERY pixel, so the only exchange is the sum of
// the per-GPU XYZAW films into the root's (Film::put under the mutex in the reference, hdrfilm.cpp:43-46).  One
// process drives one GPU; each process allocates its film with cudaMalloc, exports it as a CUDA IPC handle and opens
// its peers'.  The reduction is ONE kernel on the root that waits for the peers' films and pulls them through
// NVLink / NVSwitch with 128-bit loads, adding in rank order (deterministic), so there is no NCCL call, no host
// barrier and no staging copy on the data path:
//
//   peer  (stream order)   render -> k_film_publish:   fence.sys; ctrl.ready = epoch
//                                    k_film_wait_consumed: spin until ctrl.consumed == epoch (written by the root)
//   root  (stream order)   render -> k_film_reduce:    every block spins until all peers' ready == epoch, then
//                                    film[i] += sum_p peer_p.film[i]; the last block stores consumed = epoch into
//                                    every peer's control word
//
// Flags are polled with volatile (L1-bypassing) loads and published after __threadfence_system(); film data is read
// with ld.global.cg.  Every spin has a wall-clock bound (MSK_PEER_TIMEOUT_S, default 30 s) after which the kernel gives up and
// raises an error word that the host reports, so a crashed peer cannot hang the GPU.  The time-out decision is
// GRID-WIDE: every block records the outcome of its spin, the blocks meet at a counter, and either all of them add the
// peers' films or none does; after a time-out the peers are NOT released (their own wait then times out and reports
// too) and the film is left as the root rendered it.  msk_gpu_film_share_check reports and clears the error.
//
// One process driving several GPUs (the host plugin's `devices` property, msk_gpu_render_multi) uses the same kernels
// on pointers made visible with cudaDeviceEnablePeerAccess instead of IPC handles (msk_gpu_film_share_attach).
#include "msk_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace msk;

namespace {

constexpr size_t kCtrlBytes = 256; // control block in front of the film, keeps the film 256-byte aligned
// wall-clock bound of every device-side spin, MSK_PEER_TIMEOUT_S (default 30 s)
unsigned long long peer_timeout_ns() {
    const char *v = getenv("MSK_PEER_TIMEOUT_S");
    double sec = (v && *v) ? atof(v) : 30.0;
    if (!(sec > 0.0)) sec = 30.0;
    return (unsigned long long) (sec * 1e9);
}

struct PeerCtrl {
    uint32_t ready;     // written by the owner: epoch of the film that is complete
    uint32_t consumed;  // written by the root: epoch it has finished reading
    uint32_t error;     // a spin timed out
    uint32_t done_blocks; // root only: blocks of k_film_reduce that have finished
    uint32_t arrived;     // root only: blocks of k_film_reduce whose spin has ended (grid-wide time-out decision)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool spin_until(const volatile uint32_t *word, uint32_t epoch, unsigned long long timeout_ns) {
    const unsigned long long t0 = globaltimer_ns();
    while (*word != epoch) {
        __nanosleep(200);
        if (globaltimer_ns() - t0 > timeout_ns) return false;
    }
    return true;
}

__global__ void k_film_publish(PeerCtrl *ctrl, uint32_t epoch) {
    __threadfence_system(); // the film written by the preceding kernels of this stream is visible system-wide
    *reinterpret_cast<volatile uint32_t *>(&ctrl->ready) = epoch;
}

__global__ void k_film_wait_consumed(PeerCtrl *ctrl, uint32_t epoch, unsigned long long timeout_ns) {
    if (!spin_until(&ctrl->consumed, epoch, timeout_ns)) ctrl->error = 1;
}

struct PeerList {
    const float4 *film[15];
    PeerCtrl *ctrl[15];
    uint32_t n;
};

__global__ void __launch_bounds__(256) k_film_reduce(float4 *film, PeerCtrl *self, PeerList peers, size_t n4, uint32_t epoch,
                                                     unsigned long long timeout_ns) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        int good = 1;
        for (uint32_t p = 0; p < peers.n; ++p) good &= spin_until(&peers.ctrl[p]->ready, epoch, timeout_ns) ? 1 : 0;
        if (!good) atomicExch(&self->error, 1u);
        __threadfence();
        // every block of the grid is resident (<= 4 blocks of 256 threads per SM), so this counter is a grid barrier:
        // once all spins have ended, all blocks read the same error word
        atomicAdd(&self->arrived, 1u);
        if (!spin_until(&self->arrived, gridDim.x, timeout_ns)) atomicExch(&self->error, 1u);
        ok = *reinterpret_cast<volatile uint32_t *>(&self->error) == 0u;
        __threadfence_system(); // acquire: order the film loads below after the flag loads
    }
    __syncthreads();
    if (ok) {
        for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) {
            float4 acc = film[i];
            for (uint32_t p = 0; p < peers.n; ++p) { // rank order: the sum is deterministic
                const float4 v = __ldcg(peers.film[p] + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            film[i] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&self->done_blocks, 1u) == gridDim.x - 1) { // last block: release the peers' films (not after a time-out)
            self->done_blocks = 0; self->arrived = 0;
            if (ok)
                for (uint32_t p = 0; p < peers.n; ++p) *reinterpret_cast<volatile uint32_t *>(&peers.ctrl[p]->consumed) = epoch;
        }
    }
}

} // namespace

struct MskFilmShare {
    int device = -1;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t nfloats = 0;
    unsigned char *base = nullptr; // [PeerCtrl | pad to 256 B][film]
    std::vector<void *> peer_bases;
    bool peers_are_ipc = true;     // false: in-process peers (msk_gpu_film_share_attach), nothing to close
    uint32_t *h_error = nullptr;
    unsigned long long timeout_ns = 0;
};

// defined in msk_api.cu
extern "C" void *msk_gpu_stream(MskCtx *ctx);
int msk_ctx_device(MskCtx *ctx);
int msk_ctx_sm_count(MskCtx *ctx);

extern "C" {

int msk_gpu_film_share_create(MskCtx *ctx, size_t nfloats, MskFilmShare **out) {
    if (!ctx || !out || !nfloats || (nfloats & 3u)) return fail(MSK_ERR_ARG, "msk_gpu_film_share_create: the film size must be a positive multiple of 4 floats");
    *out = nullptr;
    MskFilmShare *s = new (std::nothrow) MskFilmShare;
    if (!s) return fail(MSK_ERR_OOM, "out of host memory");
    s->device = msk_ctx_device(ctx); s->sm_count = msk_ctx_sm_count(ctx);
    s->stream = (cudaStream_t) msk_gpu_stream(ctx);
    s->nfloats = nfloats;
    s->timeout_ns = peer_timeout_ns();
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    // CUDA loads a kernel's code at its first launch, and that load may synchronise the whole context.  A root whose
    // k_film_reduce is already spinning would then wait for a peer OF THE SAME CONTEXT (two MskCtx on one GPU) whose first
    // k_film_publish can never be loaded: force the three kernels in now.
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_film_reduce); cudaFuncGetAttributes(&fa, k_film_publish); cudaFuncGetAttributes(&fa, k_film_wait_consumed);
    cudaError_t e = cudaMalloc((void **) &s->base, kCtrlBytes + nfloats * sizeof(float)); // plain cudaMalloc: exportable
    if (e == cudaSuccess) e = cudaMemset(s->base, 0, kCtrlBytes + nfloats * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost((void **) &s->h_error, sizeof(uint32_t));
    if (prev >= 0 && prev != s->device) cudaSetDevice(prev);
    if (e != cudaSuccess) { cudaFree(s->base); delete s; return cuda_fail(e, "film share allocation", __FILE__, __LINE__); }
    *out = s;
    return MSK_OK;
}

float *msk_gpu_film_share_ptr(MskFilmShare *s) { return s ? reinterpret_cast<float *>(s->base + kCtrlBytes) : nullptr; }

int msk_gpu_film_share_export(MskFilmShare *s, MskIpcMemHandle *out) {
    if (!s || !out) return fail(MSK_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(MskIpcMemHandle), "handle size");
    cudaIpcMemHandle_t h;
    MSK_CUDA_CHECK(cudaIpcGetMemHandle(&h, s->base));
    memcpy(out, &h, sizeof(h));
    return MSK_OK;
}

int msk_gpu_film_share_open(MskFilmShare *s, const MskIpcMemHandle *peers, uint32_t npeers) {
    if (!s || (npeers && !peers)) return fail(MSK_ERR_ARG, "null argument");
    if (npeers > 15) return fail(MSK_ERR_UNSUPPORTED, "at most 16 GPUs per reduction");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    if (s->peers_are_ipc) for (void *p : s->peer_bases) cudaIpcCloseMemHandle(p);
    s->peer_bases.clear();
    s->peers_are_ipc = true;
    int rc = MSK_OK;
    for (uint32_t i = 0; i < npeers && rc == MSK_OK; ++i) {
        cudaIpcMemHandle_t h;
        memcpy(&h, &peers[i], sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaIpcOpenMemHandle (peer film)", __FILE__, __LINE__);
        else s->peer_bases.push_back(p);
    }
    if (prev >= 0 && prev != s->device) cudaSetDevice(prev);
    return rc;
}

// In-process variant of msk_gpu_film_share_open: the peers' shares live in this process (one MskCtx per device), so the root
// maps their memory with cudaDeviceEnablePeerAccess instead of IPC handles.  A peer on the root's own device (two contexts on
// one GPU: the single-GPU test of the multi-device path) needs no mapping.
int msk_gpu_film_share_attach(MskFilmShare *s, MskFilmShare *const *peers, uint32_t npeers) {
    if (!s || (npeers && !peers)) return fail(MSK_ERR_ARG, "null argument");
    if (npeers > 15) return fail(MSK_ERR_UNSUPPORTED, "at most 16 GPUs per reduction");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    if (s->peers_are_ipc) for (void *p : s->peer_bases) cudaIpcCloseMemHandle(p);
    s->peer_bases.clear();
    s->peers_are_ipc = false;
    int rc = MSK_OK;
    for (uint32_t i = 0; i < npeers && rc == MSK_OK; ++i) {
        MskFilmShare *p = peers[i];
        if (!p || p->nfloats != s->nfloats) { rc = fail(MSK_ERR_ARG, "peer film %u: null or of another size", i); break; }
        if (p->device != s->device) {
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, s->device, p->device);
            if (e != cudaSuccess || !can) { rc = fail(MSK_ERR_UNSUPPORTED, "device %d cannot access the memory of device %d (no NVLink / PCIe peer path)", s->device, p->device); break; }
            e = cudaDeviceEnablePeerAccess(p->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); break; }
        }
        s->peer_bases.push_back(p->base);
    }
    if (rc != MSK_OK) s->peer_bases.clear();
    if (prev >= 0 && prev != s->device) cudaSetDevice(prev);
    return rc;
}

// Root (is_root != 0): film += sum of the peers' films of this epoch, on the context's stream.  Other ranks: publish the
// local film for this epoch and hold the stream until the root has read it.  `epoch` must be non-zero and change from
// one reduction to the next (a step counter); every rank passes the same value.  Asynchronous; errors of the spin
// time-outs surface in msk_gpu_film_share_check.
int msk_gpu_reduce_film(MskFilmShare *s, int is_root, uint32_t epoch) {
    if (!s || !epoch) return fail(MSK_ERR_ARG, "msk_gpu_reduce_film: null share or zero epoch");
    PeerCtrl *self = reinterpret_cast<PeerCtrl *>(s->base);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    if (is_root) {
        PeerList pl{};
        pl.n = (uint32_t) s->peer_bases.size();
        for (uint32_t i = 0; i < pl.n; ++i) {
            pl.ctrl[i] = reinterpret_cast<PeerCtrl *>(s->peer_bases[i]);
            pl.film[i] = reinterpret_cast<const float4 *>((unsigned char *) s->peer_bases[i] + kCtrlBytes);
        }
        const size_t n4 = s->nfloats / 4;
        const int blocks = (int) std::min<size_t>((size_t) s->sm_count * 4, (n4 + 255) / 256);
        if (pl.n) k_film_reduce<<<blocks, 256, 0, s->stream>>>(reinterpret_cast<float4 *>(s->base + kCtrlBytes), self, pl, n4, epoch, s->timeout_ns);
    } else {
        k_film_publish<<<1, 1, 0, s->stream>>>(self, epoch);
        k_film_wait_consumed<<<1, 1, 0, s->stream>>>(self, epoch, s->timeout_ns);
    }
    cudaError_t e = cudaGetLastError();
    if (prev >= 0 && prev != s->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "film reduction launch", __FILE__, __LINE__);
    return MSK_OK;
}

// Synchronises the stream and reports a timed-out spin (a peer that never published / a root that never consumed).
int msk_gpu_film_share_check(MskFilmShare *s) {
    if (!s) return fail(MSK_ERR_ARG, "null share");
    PeerCtrl *self = reinterpret_cast<PeerCtrl *>(s->base);
    MSK_CUDA_CHECK(cudaMemcpyAsync(s->h_error, &self->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    MSK_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    if (*s->h_error) { // report once, then clear so that the share is usable again (the counters of the aborted launch too)
        MSK_CUDA_CHECK(cudaMemsetAsync(&self->error, 0, 3 * sizeof(uint32_t), s->stream));
        MSK_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        return fail(MSK_ERR_CUDA, "film reduction timed out waiting for a peer GPU; the film holds this GPU's samples only");
    }
    return MSK_OK;
}

void msk_gpu_film_share_destroy(MskFilmShare *s) {
    if (!s) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (s->peers_are_ipc) for (void *p : s->peer_bases) cudaIpcCloseMemHandle(p);
    cudaFree(s->base);
    if (s->h_error) cudaFreeHost(s->h_error);
    if (prev >= 0 && prev != s->device) cudaSetDevice(prev);
    delete s;
}

} // extern "C"
