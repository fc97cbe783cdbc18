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

int workspace_reserve_accum(spb_workspace* ws, size_t bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    if (ws->accum_bytes >= bytes && ws->accum) return SPB_OK;
    SPB_NO_GROWTH_IN_CAPTURE(stream);
    if (ws->accum) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->accum));
        ws->accum = nullptr;
        ws->accum_bytes = 0;
    }
    const size_t want = (bytes + bytes / 4 + (1u << 20) + 255) & ~(size_t)255;
    SPB_CUDA_TRY(cudaMalloc(&ws->accum, want));
    SPB_CUDA_TRY(cudaMemsetAsync(ws->accum, 0, want, stream));     // zero from here on: users clean up after themselves
    ws->accum_bytes = want;
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
    delete ws;
    return SPB_OK;
}

}  // extern "C"
