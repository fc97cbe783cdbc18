// sparray_b200/csrc/workspace.cu -- per-stream scratch: look-back state words (epoch-tagged,
// never memset between launches), a small plain scratch slab, and a payload buffer.
#include "common.cuh"

namespace spb {

// Is this stream being captured into a CUDA graph?  Buffers cannot grow then (cudaMalloc / cudaFree /
// cudaStreamSynchronize are illegal during capture): run the same call once before capturing.
static bool capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}
#define SPB_NO_GROWTH_IN_CAPTURE(stream) \
    do {                                 \
        if (capturing(stream)) return SPB_ERR_CAPTURE; \
    } while (0)

int workspace_reserve(spb_workspace* ws, size_t state_bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    const size_t bytes = kWsMiscBytes + state_bytes;
    ws->last_state_bytes = state_bytes;
    if (ws->bytes >= bytes && ws->dev) return SPB_OK;
    SPB_NO_GROWTH_IN_CAPTURE(stream);
    size_t want = (bytes + state_bytes / 2 + (1u << 20) + 255) & ~(size_t)255;
    if (ws->dev) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->dev));
        ws->dev = nullptr;
        ws->bytes = 0;
    }
    SPB_CUDA_TRY(cudaMalloc(&ws->dev, want + kWsCounterBytes));
    SPB_CUDA_TRY(cudaMemsetAsync(ws->dev, 0, want + kWsCounterBytes, stream));
    ws->bytes = want;
    // The epoch counter keeps running: the fresh buffer is all zero (= "invalid" in every epoch), while
    // the 16-byte entries of state16 are NOT reallocated here and still hold words tagged with earlier
    // epochs -- replaying 1, 2, 3... would make them look live again.
    return SPB_OK;
}

int workspace_reserve_payload(spb_workspace* ws, size_t bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    if (ws->payload_bytes >= bytes && ws->payload) return SPB_OK;
    SPB_NO_GROWTH_IN_CAPTURE(stream);
    if (ws->payload) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->payload));
        ws->payload = nullptr;
        ws->payload_bytes = 0;
    }
    const size_t want = bytes + bytes / 2 + (1u << 20);
    SPB_CUDA_TRY(cudaMalloc(&ws->payload, want));
    ws->payload_bytes = want;
    return SPB_OK;
}

int workspace_reserve_state16(spb_workspace* ws, size_t bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    ws->last_state16_bytes = bytes;
    if (ws->state16_bytes >= bytes && ws->state16) return SPB_OK;
    SPB_NO_GROWTH_IN_CAPTURE(stream);
    if (ws->state16) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->state16));
        ws->state16 = nullptr;
        ws->state16_bytes = 0;
    }
    const size_t want = bytes + bytes / 2 + (1u << 20);
    SPB_CUDA_TRY(cudaMalloc(&ws->state16, want));
    SPB_CUDA_TRY(cudaMemsetAsync(ws->state16, 0, want, stream));   // zero words read as "invalid" in every epoch
    ws->state16_bytes = want;
    return SPB_OK;
}

__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static int fill_negzero(void* p, size_t bytes, cudaStream_t stream) {
    const size_t n = bytes / 8;
    size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > (size_t)device_sm_count() * 16) blocks = (size_t)device_sm_count() * 16;
    if (blocks == 0) return SPB_OK;
    fill_u64_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<unsigned long long*>(p), n, kNegZeroBits);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

// The dense accumulators of sum(axis): float64 cells that hold -0.0 between calls, and one flag byte per cell
// (zero between calls) for the sums that mark touched cells with flags.  Users clean up after themselves.
int workspace_reserve_accum(spb_workspace* ws, size_t acc_bytes, size_t flag_bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    if (ws->accum_bytes < acc_bytes || !ws->accum) {
        SPB_NO_GROWTH_IN_CAPTURE(stream);
        if (ws->accum) {
            SPB_CUDA_TRY(cudaStreamSynchronize(stream));
            SPB_CUDA_TRY(cudaFree(ws->accum));
            ws->accum = nullptr;
            ws->accum_bytes = 0;
        }
        const size_t want = (acc_bytes + acc_bytes / 4 + (1u << 20) + 255) & ~(size_t)255;
        SPB_CUDA_TRY(cudaMalloc(&ws->accum, want));
        const int rc = fill_negzero(ws->accum, want, stream);
        if (rc) return rc;
        ws->accum_bytes = want;
    }
    if (flag_bytes && (ws->flags_bytes < flag_bytes || !ws->flags)) {
        SPB_NO_GROWTH_IN_CAPTURE(stream);
        if (ws->flags) {
            SPB_CUDA_TRY(cudaStreamSynchronize(stream));
            SPB_CUDA_TRY(cudaFree(ws->flags));
            ws->flags = nullptr;
            ws->flags_bytes = 0;
        }
        const size_t want = (flag_bytes + flag_bytes / 4 + (1u << 16) + 255) & ~(size_t)255;
        SPB_CUDA_TRY(cudaMalloc(&ws->flags, want));
        SPB_CUDA_TRY(cudaMemsetAsync(ws->flags, 0, want, stream));
        ws->flags_bytes = want;
    }
    return SPB_OK;
}

// back to the clean state after a call that failed between its accumulating and its emitting launch
int workspace_wipe_accum(spb_workspace* ws, cudaStream_t stream) {
    if (ws->accum) {
        const int rc = fill_negzero(ws->accum, ws->accum_bytes, stream);
        if (rc) return rc;
    }
    if (ws->flags) SPB_CUDA_TRY(cudaMemsetAsync(ws->flags, 0, ws->flags_bytes, stream));
    ws->accum_dirty = 0;
    return SPB_OK;
}

// next launch epoch (1 .. 2^14-1); re-zero the state words when the counter wraps.
// CUDA graphs: a captured launch carries its epoch as a baked-in kernel argument, so every replay would
// meet the words of the previous replay tagged with the very same epoch.  While the stream is being
// captured, the words the call asked for (workspace_reserve / _state16 just before) are therefore wiped by
// a memset node in front of the launch: a replayed chain is self-contained.
// wipe_in_capture = false: the caller's kernels do not read look-back words, or a producer launch of the
// same chain wipes them.
int workspace_next_epoch(spb_workspace* ws, cudaStream_t stream, uint32_t* epoch, bool wipe_in_capture) {
    if (wipe_in_capture && capturing(stream)) {
        if (ws->last_state_bytes)
            SPB_CUDA_TRY(cudaMemsetAsync((char*)ws->dev + kWsMiscBytes, 0, ws->last_state_bytes, stream));
        if (ws->state16 && ws->last_state16_bytes)
            SPB_CUDA_TRY(cudaMemsetAsync(ws->state16, 0, ws->last_state16_bytes, stream));
    }
    ws->epoch += 1;
    if (ws->epoch >= (1u << kEpochBits)) {
        SPB_CUDA_TRY(cudaMemsetAsync(ws->dev, 0, ws->bytes, stream));
        if (ws->state16) SPB_CUDA_TRY(cudaMemsetAsync(ws->state16, 0, ws->state16_bytes, stream));
        ws->epoch = 1;
    }
    *epoch = ws->epoch;
    return SPB_OK;
}

int device_sm_count() {
    static std::atomic<int> sms[64];
    const int dev = current_device() & 63;
    int v = sms[dev].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace spb

extern "C" {

int spb_workspace_create(spb_workspace** ws) {
    if (!ws) return SPB_ERR_BAD_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SPB_ERR_NO_DEVICE;
    spb_workspace* w = new spb_workspace();
    w->dev = nullptr;
    w->bytes = 0;
    w->epoch = 0;
    w->device = dev;
    w->payload = nullptr;
    w->payload_bytes = 0;
    w->state16 = nullptr;
    w->state16_bytes = 0;
    w->accum = nullptr;
    w->accum_bytes = 0;
    w->flags = nullptr;
    w->flags_bytes = 0;
    w->accum_dirty = 0;
    w->last_state_bytes = 0;
    w->last_state16_bytes = 0;
    *ws = w;
    return SPB_OK;
}

int spb_workspace_destroy(spb_workspace* ws) {
    if (!ws) return SPB_OK;
    if (ws->dev) cudaFree(ws->dev);
    if (ws->payload) cudaFree(ws->payload);
    if (ws->state16) cudaFree(ws->state16);
    if (ws->accum) cudaFree(ws->accum);
    if (ws->flags) cudaFree(ws->flags);
    delete ws;
    return SPB_OK;
}

}  // extern "C"
