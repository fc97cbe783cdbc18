/*
 * sparray_b200.h -- C ABI of libsparray_b200.so: the B200 (sm_100a) replacement
 * for the native layer of perimosocordiae/sparray's FlatSparray hot path.
 *
 * Citations `file:line` are into the reference repository.  Every entry point
 * works on raw device pointers (plain C types only, no torch types), is
 * asynchronous on `stream`, returns 0 or a cudaError_t / SPB_ERR_* code, and
 * never falls back to the CPU.  Outputs are caller-allocated at the worst-case
 * size stated per function; data-dependent output counts are written to a
 * device int64 (`*_count`) so a chain of calls needs no host synchronisation.
 *
 * The reference's own FFI for this path is the four-name seam that
 * sparray/flat.py:7-10 imports from sparray/compat.py:104-120 (implemented by
 * sparray/_merge.pyx); section 1 below is the one-to-one replacement of those
 * four names, on device buffers (`spb_*`) and on host buffers (`spb_host_*`,
 * which is what a ctypes/Cython shim inside the reference would bind -- see
 * INTEGRATION.md).  Sections 2-6 replace the numpy/scipy work that flat.py
 * does around the seam (fused so the lut/mask arrays never exist).
 */
#ifndef SPARRAY_B200_H_
#define SPARRAY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* spb_stream_t;   /* == cudaStream_t */

/* value dtypes (numpy names) */
enum spb_dtype {
    SPB_F32 = 0, SPB_F64 = 1, SPB_I32 = 2, SPB_I64 = 3, SPB_U8 = 4, SPB_BOOL = 5, SPB_I8 = 6, SPB_I16 = 7,
    /* conversion-only dtypes: spb_cast accepts them on either side; every other entry point works on the
     * dtypes above.  The host widens them for arithmetic (float16 -> float32, uint16 -> int32, uint32 -> int64)
     * and passes them as same-sized integers to the kernels that only move values. */
    SPB_F16 = 8, SPB_U16 = 9, SPB_U32 = 10, SPB_U64 = 11,
    SPB_DTYPE_COUNT = 12
};

/* binary value ops fused into the set merges (numpy ufunc semantics incl. NaN / signed zero) */
enum spb_binop {
    SPB_ADD = 0, SPB_SUB = 1, SPB_MUL = 2, SPB_MINIMUM = 3, SPB_MAXIMUM = 4,
    SPB_LT = 5, SPB_LE = 6, SPB_GT = 7, SPB_GE = 8, SPB_EQ = 9, SPB_NE = 10,   /* -> bool output */
    SPB_TAKE_A = 11, SPB_TAKE_B = 12,
    SPB_OVERRIDE = 13,   /* union only: b where b is stored, else a -- _setitem_flatidx, flat.py:395-407 */
    SPB_DIVIDE = 14, SPB_FLOOR_DIVIDE = 15,   /* spb_binary_map only (dense operand, flat.py:444-449,534-536) */
    SPB_BINOP_COUNT = 16
};

/* unary / scalar value maps: flat.py:580-581,677-739, compat.py:96-102 */
enum spb_unop {
    SPB_U_COPY = 0, SPB_U_NEG, SPB_U_ABS, SPB_U_MUL_S, SPB_U_DIV_S, SPB_U_FLOORDIV_S, SPB_U_POW_S,
    SPB_U_MIN_S, SPB_U_MAX_S, SPB_U_LT_S, SPB_U_LE_S, SPB_U_GT_S, SPB_U_GE_S, SPB_U_EQ_S, SPB_U_NE_S,
    SPB_U_SIN, SPB_U_TAN, SPB_U_ARCSIN, SPB_U_ARCTAN, SPB_U_SINH, SPB_U_TANH, SPB_U_ARCSINH, SPB_U_ARCTANH,
    SPB_U_RINT, SPB_U_SIGN, SPB_U_EXPM1, SPB_U_LOG1P, SPB_U_DEG2RAD, SPB_U_RAD2DEG, SPB_U_FLOOR, SPB_U_CEIL,
    SPB_U_TRUNC, SPB_U_SQRT, SPB_U_RDIV_S,
    SPB_UNOP_COUNT
};

enum spb_error {
    SPB_OK = 0,
    SPB_ERR_BAD_ARG = 10001,        /* null pointer, negative size, unknown enum */
    SPB_ERR_UNSUPPORTED = 10002,    /* dtype/op combination not instantiated */
    SPB_ERR_NO_DEVICE = 10003,      /* no CUDA device: there is no CPU fallback */
    SPB_ERR_TOO_LARGE = 10004,
    SPB_ERR_CAPTURE = 10005         /* scratch would have to grow while the stream is captured into a CUDA graph:
                                       run the same call once before capturing */
};

/* ---- 0. library / workspace ------------------------------------------------ */

const char* spb_version(void);
const char* spb_error_string(int code);
int spb_device_count(void);                 /* 0 if no usable CUDA device */
unsigned long long spb_launch_count(void);  /* kernels launched by this library so far */
uint64_t spb_debug_fastdiv(uint64_t n, uint64_t d);   /* host copy of the kernels' exact n / d (n < 2^63) */

/* Scratch for tile look-back state etc.  One per stream in flight; grows on
 * demand.  State words are epoch-tagged, so no memset between launches.
 * CUDA graphs: every entry point may be captured (cudaStreamBeginCapture / torch.cuda.graph) once the same
 * call has run eagerly, so that the scratch has its size; a captured launch wipes the state words it uses
 * with a memset node of its own (an epoch baked into a graph cannot tell one replay from the next). */
typedef struct spb_workspace spb_workspace;
int spb_workspace_create(spb_workspace** ws);
int spb_workspace_destroy(spb_workspace* ws);

/* ---- 1. the compat seam: compat.py:104-120 / _merge.pyx -------------------- */

/* union1d_sorted (_merge.pyx:92-140).  a, b sorted, duplicate-free int64.
 * out (na+nb), lut (na+nb: 0 a-only / 1 b-only / 2 both), a_only (na), b_only (nb)
 * may each be NULL except out.  out_count: device int64. */
int spb_union1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                       int64_t* out, uint8_t* lut, uint8_t* a_only, uint8_t* b_only,
                       int64_t* out_count, spb_workspace* ws, spb_stream_t stream);

/* intersect1d_sorted (_merge.pyx:81-89,145-157).  out, a_inds, b_inds have
 * min(na,nb) slots; a_inds/b_inds may be NULL. */
int spb_intersect1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                           int64_t* out, int64_t* a_inds, int64_t* b_inds,
                           int64_t* out_count, spb_workspace* ws, spb_stream_t stream);

/* combine_ranges (_merge.pyx:7-66).  ranges: HOST int64[ndim][3] rows of
 * (start, stop, step); shape: HOST int64[ndim]; out: DEVICE int64[result_size]. */
int spb_combine_ranges(const int64_t* ranges_host, const int64_t* shape_host, int ndim,
                       int64_t* out, int64_t result_size, int inner, spb_stream_t stream);

/* len_range (_merge.pyx:70-78); pure host arithmetic. */
int64_t spb_len_range(int64_t start, int64_t stop, int64_t step);

/* Host-buffer forms of the same four names: copy in, run the kernels above,
 * copy out, synchronise.  *out_count is a HOST int64. */
int spb_host_union1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                            int64_t* out, uint8_t* lut, uint8_t* a_only, uint8_t* b_only,
                            int64_t* out_count);
int spb_host_intersect1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                                int64_t* out, int64_t* a_inds, int64_t* b_inds, int64_t* out_count);
int spb_host_combine_ranges(const int64_t* ranges, const int64_t* shape, int ndim,
                            int64_t* out, int64_t result_size, int inner);

/* ---- 2. fused sparse (op) sparse: flat.py:409-434 -------------------------- */

/* _pairwise_sparray (flat.py:409-422): out = op(a or 0, b or 0) on the union;
 * out_idx/out_val need na+nb slots.  Both value arrays have dtype `dtype`
 * (the host promotes first, flat.py:415); comparisons write SPB_BOOL bytes. */
int spb_union_binop(int op, int dtype,
                    const int64_t* a_idx, const void* a_val, int64_t na,
                    const int64_t* b_idx, const void* b_val, int64_t nb,
                    int64_t* out_idx, void* out_val, int64_t* out_count,
                    spb_workspace* ws, spb_stream_t stream);

/* _pairwise_sparray_fixed_zero (flat.py:424-434): out = op(a, b) on the
 * intersection; min(na,nb) slots. */
int spb_intersect_binop(int op, int dtype,
                        const int64_t* a_idx, const void* a_val, int64_t na,
                        const int64_t* b_idx, const void* b_val, int64_t nb,
                        int64_t* out_idx, void* out_val, int64_t* out_count,
                        spb_workspace* ws, spb_stream_t stream);

/* ---- 3. value maps, conversion, lookup: flat.py:48-54,74-77,262-267,580-581,677-739 ---- */

/* `_with_data(f(data))`: out[i] = f(in[i]) or f(in[i], scalar).  The scalar is
 * read from scalar_f for float dtypes and scalar_i for integer dtypes.
 * Comparisons (SPB_U_LT_S..NE_S) write SPB_BOOL bytes; everything else keeps
 * `dtype` (the host promotes first, following numpy's scalar rules). */
int spb_unary_map(int op, int dtype, const void* in, void* out, int64_t n,
                  double scalar_f, int64_t scalar_i, spb_stream_t stream);

/* out[i] = a[i] (op) b[i] on two aligned value arrays of dtype `dtype` (comparisons -> SPB_BOOL):
 * the dense (op) sparse helpers _pairwise_dense2dense / _pairwise_dense2sparse (flat.py:436-449)
 * once the dense operand has been gathered at the stored indices with spb_gather. */
int spb_binary_map(int op, int dtype, const void* a, const void* b, void* out, int64_t n, spb_stream_t stream);

/* astype (flat.py:713-714) */
int spb_cast(int dtype_in, int dtype_out, const void* in, void* out, int64_t n, spb_stream_t stream);

/* complex64 / complex128 values (numpy's interleaved re, im storage) <-> two real arrays of real_dtype (SPB_F32 /
 * SPB_F64).  The only complex-aware entry points: the host composes every complex operation of FlatSparray from the
 * real kernels on the two parts, which always share one index array (flat.py:19-42 accepts any numpy dtype). */
int spb_complex_split(int real_dtype, const void* c, int64_t n, void* re, void* im, spb_stream_t stream);
int spb_complex_join(int real_dtype, const void* re, const void* im, int64_t n, void* c, spb_stream_t stream);

/* toarray (flat.py:74-77): dense[idx[i]] = val[i]; dense is pre-zeroed by the caller */
int spb_scatter_dense(int dtype, const int64_t* idx, const void* val, int64_t n, void* dense,
                      spb_stream_t stream);

/* out[i] = src[pos[i]] (flat.py:389,431-432) */
int spb_gather(int dtype, const void* src, const int64_t* pos, int64_t n, void* out, spb_stream_t stream);

/* np.searchsorted(sorted, query) (flat.py:264): out_pos[q] = lower bound;
 * found[q] (optional) = 1 when sorted[out_pos[q]] == query[q] */
int spb_lower_bound(const int64_t* sorted, int64_t n, const int64_t* query, int64_t nq,
                    int64_t* out_pos, uint8_t* found, spb_stream_t stream);

/* from_ndarray (flat.py:48-54): flat indices and values of the non-zero cells
 * of a C-contiguous dense array; n slots each. */
int spb_select_nonzero(int dtype, const void* dense, int64_t n, int64_t* out_idx, void* out_val,
                       int64_t* out_count, spb_workspace* ws, spb_stream_t stream);

/* data.sum() (flat.py:610,616): *out is a device double for float dtypes and a
 * device int64 for integer / bool dtypes (float32 accumulates in float64). */
int spb_sum_all(int dtype, const void* val, int64_t n, void* out, spb_workspace* ws, spb_stream_t stream);

/* Order-preserving compaction of (idx, val) pairs by a byte mask (flat.py:94: indices[data != 0]) */
int spb_select_flagged(int dtype, const int64_t* idx, const void* val, const uint8_t* keep, int64_t n,
                       int64_t* out_idx, void* out_val, int64_t* out_count,
                       spb_workspace* ws, spb_stream_t stream);

/* data.min() / data.max() (flat.py:719-725), NaN-propagating; *out has dtype `dtype` */
int spb_min_max(int dtype, const void* vals, int64_t n, int is_max, void* out,
                spb_workspace* ws, spb_stream_t stream);

/* ---- 4. reductions over sorted keys: flat.py:34-35, 538-568, 607-626 -------- */

/* Group-by-key float64 sum over a SORTED key stream; the segment key is
 * keys[i] / key_divisor (divisor 1: the keys themselves).  This is
 * np.unique(flat_inds) + np.bincount(inverse, weights) of sum(axis)
 * (flat.py:620-625; float64 output always) without the sort when
 * axis is the last one, and the duplicate-summing of the ctor (flat.py:34-35).
 * out_keys / out_sums: n slots.  Values of any dtype are accumulated in float64. */
int spb_reduce_by_key(int dtype, const int64_t* keys, const void* vals, int64_t n, int64_t key_divisor,
                      int64_t* out_keys, double* out_sums, int64_t* out_count,
                      spb_workspace* ws, spb_stream_t stream);

/* y[r] = sum_{i: idx[i] / ncols == r} vals[i] * x[idx[i] % ncols], r < nrows:
 * a.dot(x) for a dense vector x without densifying `a` (flat.py:538-568; the
 * reference's dense route, :568, is O(prod(shape))).  x has the dtype of vals
 * (SPB_F32 / SPB_F64); y is float64[nrows] and is fully overwritten. */
int spb_spmv(int dtype, const int64_t* idx, const void* vals, int64_t n, int64_t ncols, int64_t nrows,
             const void* x, double* y, spb_workspace* ws, spb_stream_t stream);

/* Sparse x sparse matrix product C = A (m x k) . B (k x n) on sorted flat indices -- the
 * reshape -> COO -> scipy CSR matmul -> from_spmatrix detour of flat.py:545-561 -- as expand / sort /
 * compress.  b_row_ptr[j] = lower bound of j * n in B's indices (k + 1 entries, spb_lower_bound).
 *   spb_spgemm_rowlen      len[e] = number of stored elements in row (a_idx[e] % k) of B
 *   spb_exclusive_scan_i64 out[e] = len[0] + ... + len[e-1]; *total (device int64) = the sum
 *   spb_spgemm_expand      one (i * n + c, a * b) pair per product, in A order; nprod = *total
 * The pairs are then made canonical with spb_rekey_sort + spb_reduce_by_key like any unsorted
 * constructor input (flat.py:32-35: duplicates are summed in float64). */
int spb_spgemm_rowlen(const int64_t* a_idx, int64_t na, int64_t k, const int64_t* b_row_ptr, int64_t* len,
                      spb_stream_t stream);
int spb_exclusive_scan_i64(const int64_t* in, int64_t n, int64_t* out, int64_t* total, spb_workspace* ws,
                           spb_stream_t stream);
int spb_spgemm_expand(int dtype, const int64_t* a_idx, const void* a_val, int64_t na, int64_t k,
                      const int64_t* offsets, const int64_t* b_row_ptr, const int64_t* b_idx, const void* b_val,
                      int64_t n, int64_t nprod, int64_t* out_idx, void* out_val, spb_stream_t stream);

/* ---- 5. re-key + stable radix sort: flat.py:32-35, 97-112, 617-620 ----------- */

/* new_key = sum_g coord_g(idx) * out_strides[g], where coord_g unravels idx by
 * the C-order strides in_strides[g] of (merged) axis group g in the old shape
 * (outermost first, in_strides[ndim-1] == 1; out_strides[g] == 0 drops the
 * group; ndim == 0 keeps the keys).  The (new_key, val) pairs are then stably
 * LSD-radix-sorted on new_key / sort_divisor, which must be < 2^sort_bits, in digits of at most
 * 8 bits -- the host strips trailing axes that already arrive in order
 * (sparray_b200/_plan.py).  This is np.unravel_index + np.ravel_multi_index +
 * the argsort inside np.unique of transpose (flat.py:109-112,34) and of
 * sum(axis) (flat.py:617-620) in (1 + passes) sweeps with no index temporaries.
 * out_*, tmp_*: n slots each (tmp_* only needed when sort_bits > 8); val may be
 * NULL (keys only).  Inputs are not modified. */
int spb_rekey_sort(int dtype, const int64_t* idx, const void* val, int64_t n,
                   int ndim, const int64_t* in_strides_host, const int64_t* out_strides_host,
                   int64_t sort_divisor, int sort_bits,
                   int64_t* out_idx, void* out_val, int64_t* tmp_idx, void* tmp_val,
                   spb_workspace* ws, spb_stream_t stream);

/* sum(axis != last) when prod(new_shape) (= cells) is small enough to hold densely: no sort.
 * Every element adds into acc[new_key] (float64 atomics; cells doubles + one flag byte per cell live in
 * the workspace, all-zero between calls: the emitting pass zeroes what it reads), then the touched
 * cells are emitted in key order (explicit zeros survive).  Strides as in spb_rekey_sort.
 * out_idx / out_sums: min(n, cells) slots.  The order of the additions is not fixed, so sums may
 * differ in the last bits from run to run.
 * n_dev (may be NULL): device int64 holding the real element count when the producer's count has
 * not been read back (the out_count of a binop on the same stream); n is then its upper bound. */
int spb_axis_sum_dense(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev,
                       int ndim, const int64_t* in_strides_host, const int64_t* out_strides_host,
                       int64_t cells, int64_t* out_idx, double* out_sums, int64_t* out_count,
                       spb_workspace* ws, spb_stream_t stream);

/* spb_axis_sum_dense restricted to the cells [cell_lo, cell_hi) of the reduced array: accumulators are
 * only held for that window (a range shard of a sharded array can only touch the rows of its own flat
 * range).  Elements whose cell falls outside the window are ignored.  out_idx / out_sums:
 * min(n, cell_hi - cell_lo) slots; emitted keys are absolute cell numbers. */
int spb_axis_sum_window(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev,
                        int ndim, const int64_t* in_strides_host, const int64_t* out_strides_host,
                        int64_t cell_lo, int64_t cell_hi, int64_t* out_idx, double* out_sums, int64_t* out_count,
                        spb_workspace* ws, spb_stream_t stream);

/* The two halves of spb_axis_sum_dense on caller-owned buffers, for the sharded axis sum of
 * SURVEY.md section 8(e) (dense per-GPU accumulators + all-reduce): acc = cells doubles padded to
 * a multiple of 16 and 16-byte aligned, touched = cells flag bytes padded to a multiple of 64 and 16-byte
 * aligned, both zeroed by the caller.  spb_emit_touched emits the touched cells of [cell_lo, cell_hi) in key order (absolute cell
 * numbers; cell_lo a multiple of 64; out_idx / out_sums: cell_hi - cell_lo slots) and zeroes them. */
int spb_axis_accumulate(int dtype, const int64_t* idx, const void* val, int64_t n,
                        int ndim, const int64_t* in_strides_host, const int64_t* out_strides_host,
                        int64_t cells, double* acc, uint8_t* touched, spb_stream_t stream);
int spb_emit_touched(double* acc, uint8_t* touched, int64_t cell_lo, int64_t cell_hi,
                     int64_t* out_idx, double* out_sums, int64_t* out_count,
                     spb_workspace* ws, spb_stream_t stream);

/* sum(axis) in one streaming pass, no dense accumulator array and no sort (csrc/slab_sum.cu): applies when
 * the elements feeding a bounded set of output cells are contiguous in the sorted input -- the reduced axis
 * is the last one, or the axes behind it span at most 4096 cells -- and the array is not extremely sparse.
 * total_cells = prod(shape); slab_cells = shape[axis] * suffix; suffix = prod(shape[axis+1:]).
 * spb_axis_sum_slab_ok answers (1 / 0) whether spb_axis_sum_slab accepts these sizes (it returns
 * SPB_ERR_UNSUPPORTED otherwise).  out_idx / out_sums: min(n, total_cells / shape[axis]) slots; float64 sums,
 * keys in the reduced shape, in order; elements whose flat index is >= total_cells are ignored. */
int spb_axis_sum_slab_ok(int64_t n, int64_t total_cells, int64_t slab_cells, int64_t suffix);
int spb_axis_sum_slab(int dtype, const int64_t* idx, const void* val, int64_t n, int64_t total_cells,
                      int64_t slab_cells, int64_t suffix, int64_t* out_idx, double* out_sums,
                      int64_t* out_count, spb_workspace* ws, spb_stream_t stream);

/* ---- 6. __getitem__: flat.py:254-325, 366-393 ---------------------------------- */

/* ints + slices (flat.py:270-274) without materialising the box: keeps the stored
 * elements whose coordinates satisfy start <= c < stop, (c - start) % step == 0 on
 * every axis (ranges: HOST int64[ndim][3], steps >= 1; shape: HOST int64[ndim]) and
 * writes their C-order position inside the box.  n slots of output. */
int spb_getitem_box(int dtype, const int64_t* idx, const void* val, int64_t n,
                    int ndim, const int64_t* shape_host, const int64_t* ranges_host,
                    int64_t* out_idx, void* out_val, int64_t* out_count,
                    spb_workspace* ws, spb_stream_t stream);

/* out[c] = sum_j offsets_j[coord_j(c)] for c in the C-order product of `ndim`
 * lists of lengths lens[j] (HOST) stored back to back in `offsets` (DEVICE):
 * the np.add.outer chain of flat.py:317-323. */
int spb_outer_sum(const int64_t* offsets, int ndim, const int64_t* lens_host, int64_t* out, int64_t n,
                  spb_stream_t stream);

/* _getitem_flatidx (flat.py:383-393): for every query cell j whose flat index
 * query[j] is stored, emit (j, val[position]) in query order.  nq slots. */
int spb_lookup_compact(int dtype, const int64_t* sorted, const void* val, int64_t n,
                       const int64_t* query, int64_t nq, int64_t* out_idx, void* out_val,
                       int64_t* out_count, spb_workspace* ws, spb_stream_t stream);

/* ---- 7. sparse broadcasting: flat.py:451-472 ------------------------------------ */

/* Replicates every stored element along the axes where in_shape is 1 and out_shape is larger
 * (both HOST int64[ndim], in_shape already left-padded with ones): out_idx/out_val get
 * n * prod(broadcast dims) pairs with flat indices in out_shape, NOT sorted (follow with
 * spb_rekey_sort).  The reference densifies here (flat.py:470-472, "TODO: fix this hack"). */
int spb_broadcast_expand(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim,
                         const int64_t* in_shape_host, const int64_t* out_shape_host,
                         int64_t* out_idx, void* out_val, spb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARRAY_B200_H_ */
