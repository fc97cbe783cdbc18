// sparray_b200/csrc/misc.cu -- library info, combine_ranges / len_range
// (_merge.pyx:7-78) and the host-buffer forms of the compat seam.
#include "common.cuh"

#include <string.h>

namespace spb {

std::atomic<unsigned long long> g_launches{0};

constexpr int kMaxDims = 32;

struct RangeParams {
    int64_t start[kMaxDims];    // already multiplied by the axis stride
    int64_t step[kMaxDims];
    int64_t len[kMaxDims];
    FastDiv div[kMaxDims];      // divide by the C-order stride of the *box*
    int ndim;
    int inner;
};

// out[c] = sum_axis (start_ax + step_ax * coord_ax(c)) -- coord by unravelling c in
// the box shape (outer), or coord_ax = c for every non-broadcast axis (inner).
__global__ void __launch_bounds__(256) combine_ranges_kernel(const RangeParams p, int64_t* __restrict__ out, int64_t n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        int64_t acc = 0;
        if (p.inner) {
            for (int ax = 0; ax < p.ndim; ++ax) acc += p.start[ax] + (p.len[ax] == 1 ? 0 : p.step[ax] * c);
        } else {
            uint64_t rem = (uint64_t)c;
            for (int ax = 0; ax < p.ndim; ++ax) {
                uint64_t q = fastdiv(rem, p.div[ax]);
                rem -= q * p.div[ax].d;
                acc += p.start[ax] + p.step[ax] * (int64_t)q;
            }
        }
        out[c] = acc;
    }
}

}  // namespace spb

using namespace spb;

extern "C" {

const char* spb_version(void) { return "sparray_b200 0.1 (sm_100a)"; }

const char* spb_error_string(int code) {
    switch (code) {
        case SPB_OK: return "ok";
        case SPB_ERR_BAD_ARG: return "bad argument";
        case SPB_ERR_UNSUPPORTED: return "dtype/op combination not supported by the CUDA kernels";
        case SPB_ERR_NO_DEVICE: return "no CUDA device (sparray_b200 has no CPU fallback)";
        case SPB_ERR_TOO_LARGE: return "problem too large for one launch";
        case SPB_ERR_CAPTURE: return "scratch must grow but the stream is being captured: run the call once before capturing";
        default: return cudaGetErrorString((cudaError_t)code);
    }
}

unsigned long long spb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// host evaluation of the device-side exact division (same magic numbers): for the CPU tests
uint64_t spb_debug_fastdiv(uint64_t n, uint64_t d) {
    const FastDiv f = make_fastdiv(d);
    if (f.magic == 0) return n >> f.shift;
    return (uint64_t)(((unsigned __int128)n * f.magic) >> 64) >> f.shift;
}

int spb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int64_t spb_len_range(int64_t start, int64_t stop, int64_t step) {
    if (step > 0) return start < stop ? (stop - start - 1) / step + 1 : 0;
    if (step < 0) return stop < start ? (start - stop - 1) / (-step) + 1 : 0;
    return 0;
}

int spb_combine_ranges(const int64_t* ranges, const int64_t* shape, int ndim, int64_t* out, int64_t result_size,
                       int inner, spb_stream_t stream) {
    if (ndim < 0 || ndim > kMaxDims || result_size < 0 || (ndim && (!ranges || !shape))) return SPB_ERR_BAD_ARG;
    if (result_size == 0) return SPB_OK;
    if (!out) return SPB_ERR_BAD_ARG;
    RangeParams p;
    memset(&p, 0, sizeof(p));
    p.ndim = ndim;
    p.inner = inner ? 1 : 0;
    int64_t stride = 1;
    int64_t strides[kMaxDims];
    for (int ax = ndim - 1; ax >= 0; --ax) {
        strides[ax] = stride;
        stride *= shape[ax];
    }
    for (int ax = 0; ax < ndim; ++ax) {
        p.start[ax] = ranges[3 * ax] * strides[ax];
        p.step[ax] = ranges[3 * ax + 2] * strides[ax];
        p.len[ax] = spb_len_range(ranges[3 * ax], ranges[3 * ax + 1], ranges[3 * ax + 2]);
    }
    int64_t box = 1;
    for (int ax = ndim - 1; ax >= 0; --ax) {
        p.div[ax] = make_fastdiv((uint64_t)(box > 0 ? box : 1));
        box *= p.len[ax];
    }
    int64_t blocks = ceil_div(result_size, 256);
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    combine_ranges_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, out, result_size);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

// ---- host-buffer forms (what a shim inside the reference's compat.py would bind) ----------

static int host_setup(spb_workspace** ws) {
    if (spb_device_count() == 0) return SPB_ERR_NO_DEVICE;
    return spb_workspace_create(ws);
}

#define SPB_HOST_TRY(expr)       \
    do {                         \
        rc = (int)(expr);        \
        if (rc) goto done;       \
    } while (0)

int spb_host_union1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* out, uint8_t* lut,
                            uint8_t* a_only, uint8_t* b_only, int64_t* out_count) {
    if (na < 0 || nb < 0 || !out_count) return SPB_ERR_BAD_ARG;
    spb_workspace* ws = nullptr;
    int rc = host_setup(&ws);
    if (rc) return rc;
    char* d = nullptr;
    const int64_t n = na + nb;
    // layout: a | b | out | count | lut | a_only | b_only
    size_t off_b = (size_t)na * 8, off_out = off_b + (size_t)nb * 8, off_cnt = off_out + (size_t)n * 8;
    size_t off_lut = off_cnt + 8, off_ao = off_lut + (size_t)n, off_bo = off_ao + (size_t)na;
    size_t total = off_bo + (size_t)nb + 16;
    int64_t cnt = 0;
    SPB_HOST_TRY(cudaMalloc(&d, total));
    if (na) SPB_HOST_TRY(cudaMemcpyAsync(d, a, (size_t)na * 8, cudaMemcpyHostToDevice, 0));
    if (nb) SPB_HOST_TRY(cudaMemcpyAsync(d + off_b, b, (size_t)nb * 8, cudaMemcpyHostToDevice, 0));
    SPB_HOST_TRY(spb_union1d_sorted((int64_t*)d, na, (int64_t*)(d + off_b), nb, (int64_t*)(d + off_out),
                                    lut ? (uint8_t*)(d + off_lut) : nullptr, a_only ? (uint8_t*)(d + off_ao) : nullptr,
                                    b_only ? (uint8_t*)(d + off_bo) : nullptr, (int64_t*)(d + off_cnt), ws, 0));
    SPB_HOST_TRY(cudaMemcpy(&cnt, d + off_cnt, 8, cudaMemcpyDeviceToHost));
    if (cnt && out) SPB_HOST_TRY(cudaMemcpy(out, d + off_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    if (cnt && lut) SPB_HOST_TRY(cudaMemcpy(lut, d + off_lut, (size_t)cnt, cudaMemcpyDeviceToHost));
    if (na && a_only) SPB_HOST_TRY(cudaMemcpy(a_only, d + off_ao, (size_t)na, cudaMemcpyDeviceToHost));
    if (nb && b_only) SPB_HOST_TRY(cudaMemcpy(b_only, d + off_bo, (size_t)nb, cudaMemcpyDeviceToHost));
    *out_count = cnt;
done:
    if (d) cudaFree(d);
    spb_workspace_destroy(ws);
    return rc;
}

int spb_host_intersect1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* out,
                                int64_t* a_inds, int64_t* b_inds, int64_t* out_count) {
    if (na < 0 || nb < 0 || !out_count) return SPB_ERR_BAD_ARG;
    spb_workspace* ws = nullptr;
    int rc = host_setup(&ws);
    if (rc) return rc;
    char* d = nullptr;
    const int64_t m = na < nb ? na : nb;
    size_t off_b = (size_t)na * 8, off_out = off_b + (size_t)nb * 8, off_ai = off_out + (size_t)m * 8;
    size_t off_bi = off_ai + (size_t)m * 8, off_cnt = off_bi + (size_t)m * 8;
    size_t total = off_cnt + 16;
    int64_t cnt = 0;
    SPB_HOST_TRY(cudaMalloc(&d, total));
    if (na) SPB_HOST_TRY(cudaMemcpyAsync(d, a, (size_t)na * 8, cudaMemcpyHostToDevice, 0));
    if (nb) SPB_HOST_TRY(cudaMemcpyAsync(d + off_b, b, (size_t)nb * 8, cudaMemcpyHostToDevice, 0));
    SPB_HOST_TRY(spb_intersect1d_sorted((int64_t*)d, na, (int64_t*)(d + off_b), nb, (int64_t*)(d + off_out),
                                        a_inds ? (int64_t*)(d + off_ai) : nullptr,
                                        b_inds ? (int64_t*)(d + off_bi) : nullptr, (int64_t*)(d + off_cnt), ws, 0));
    SPB_HOST_TRY(cudaMemcpy(&cnt, d + off_cnt, 8, cudaMemcpyDeviceToHost));
    if (cnt && out) SPB_HOST_TRY(cudaMemcpy(out, d + off_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    if (cnt && a_inds) SPB_HOST_TRY(cudaMemcpy(a_inds, d + off_ai, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    if (cnt && b_inds) SPB_HOST_TRY(cudaMemcpy(b_inds, d + off_bi, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    *out_count = cnt;
done:
    if (d) cudaFree(d);
    spb_workspace_destroy(ws);
    return rc;
}

int spb_host_combine_ranges(const int64_t* ranges, const int64_t* shape, int ndim, int64_t* out,
                            int64_t result_size, int inner) {
    if (result_size < 0) return SPB_ERR_BAD_ARG;
    if (result_size == 0) return SPB_OK;
    if (spb_device_count() == 0) return SPB_ERR_NO_DEVICE;
    int64_t* d = nullptr;
    int rc = 0;
    SPB_HOST_TRY(cudaMalloc(&d, (size_t)result_size * 8));
    SPB_HOST_TRY(spb_combine_ranges(ranges, shape, ndim, d, result_size, inner, 0));
    SPB_HOST_TRY(cudaMemcpy(out, d, (size_t)result_size * 8, cudaMemcpyDeviceToHost));
done:
    if (d) cudaFree(d);
    return rc;
}

}  // extern "C"
