/*
 * scedar_b200 -- C ABI of the B200 (sm_100a) distance-matrix / kNN path.
 *
 * This is the drop-in boundary for the ONE hot path of scedar
 * (TaylorResearchLab/scedar v0.2.0): the dense cell x cell distance matrix
 * behind scedar.eda.SampleDistanceMatrix and the row-wise k-nearest-neighbour
 * selection that consumes it.  The reference has no FFI of its own (it is pure
 * Python over NumPy / SciPy / scikit-learn); each entry point below names the
 * reference call site (file:line under scedar/) whose arithmetic it replaces.
 * INTEGRATION.md shows the ctypes binding a scedar maintainer would add.
 *
 * Conventions
 *  - Plain C: pointers and sizes only, no torch / C++ types.
 *  - Every data pointer is a DEVICE pointer unless the function name ends in
 *    _host.  `stream` is a cudaStream_t passed as void* (NULL = default
 *    stream).  Device-pointer calls are asynchronous on that stream.
 *  - The caller owns every input and output buffer.  The library allocates
 *    only scratch, stream-ordered (cudaMallocAsync) and released before return,
 *    plus the working copy inside a scedar_cells_t, released by
 *    scedar_cells_free.  Inputs are never modified.
 *  - All matrices are row-major float64; indices out are int64 (NumPy's
 *    argsort dtype); CSR indices are int32, indptr int32 or int64.
 *  - Return value: 0 on success, a SCEDAR_ERR_* code otherwise; nothing throws
 *    across the ABI.  scedar_last_error() gives the message for the calling
 *    thread's last failure.
 *  - There is no CPU fallback anywhere behind this ABI.
 *  - Inputs must be finite.  NaN / inf propagate into D as IEEE arithmetic
 *    dictates, but the selection kernels order by (distance, index) with
 *    ordinary comparisons, so a NaN distance is never selected (NumPy's sort
 *    would place it last); rows that consist of NaN yield index -1.
 */
#ifndef SCEDAR_B200_H
#define SCEDAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCEDAR_ABI_VERSION 2

/* metric: SampleDistanceMatrix(metric=...), scedar/eda/sdm.py:1211-1217 */
enum {
  SCEDAR_METRIC_COSINE = 0,      /* cosine_pdist,      sdm.py:1223-1266 */
  SCEDAR_METRIC_CORRELATION = 1, /* correlation_pdist, sdm.py:1268-1287 */
  SCEDAR_METRIC_EUCLIDEAN = 2    /* sklearn pairwise_distances, sdm.py:1216 */
};

/* precision of the Gram contraction */
enum {
  SCEDAR_PRECISION_FP64 = 0, /* float64 tensor-core (DMMA) path, <=1e-10 rel */
  SCEDAR_PRECISION_FAST = 1, /* 3xFP16 tcgen05 path: D to a stated tolerance
                                (1e-5 cosine/correlation; 2e-6 max|x|^2 on d^2
                                euclidean), kNN exact (FP64 re-rank)         */
  SCEDAR_PRECISION_FP64_I8 = 2,   /* float64-grade Gram as an error-free split
                                on the INT8 tensor pipe (tcgen05 kind::i8, 7
                                byte slices, exact INT32 accumulation, float64
                                epilogue): <=1e-10 rel like FP64, dot products
                                of integer data exact; rows longer than 4096
                                features are contracted in K chunks           */
  SCEDAR_PRECISION_FP64_I8S6 = 3  /* the same with 6 slices (21 instead of 28
                                products): |dD| <= 1e-11 of max|x|^2-scaled
                                products, see gram_oz.cu                      */
};

/* which of the k+1 nearest is dropped */
enum {
  SCEDAR_EXCLUDE_FIRST = 0, /* rank 0 of the stable sort: s_knn_ind_lut,
                               sdm.py:764 (_col_argsorted_d[1:k+1])          */
  SCEDAR_EXCLUDE_SELF = 1   /* the row itself, else rank 0: sklearn
                               kneighbors(None), sdm.py:1078-1079            */
};

enum {
  SCEDAR_OK = 0,
  SCEDAR_ERR_ARG = 1,         /* bad argument (Python side: ValueError)   */
  SCEDAR_ERR_CUDA = 2,        /* a CUDA call failed                       */
  SCEDAR_ERR_UNSUPPORTED = 3, /* valid request outside this build's range */
  SCEDAR_ERR_NOMEM = 4        /* device scratch allocation failed         */
};

/* flags written by scedar_num_correct_dist_mat */
#define SCEDAR_NCDM_BAD_DIAG 1 /* some |diag| > 1e-5           (sdm.py:196-203) */
#define SCEDAR_NCDM_ASYMMETRIC 2 /* triangles differ, rtol 1e-3 (sdm.py:205-213) */

/* flags OR-ed into *flags_out by scedar_squareform_rows: what scipy's
 * squareform(checks=True) -> is_valid_dm rejects with a ValueError */
#define SCEDAR_SQF_NAN 1      /* a NaN above the diagonal (D == D.T fails)  */
#define SCEDAR_SQF_BAD_DIAG 2 /* a diagonal entry that is not exactly zero  */
/* flag OR-ed into *flags_out by scedar_d_gather_block */
#define SCEDAR_GATHER_BAD_INDEX 1 /* a selected row / column is out of range */

int scedar_abi_version(void);
const char* scedar_last_error(void);

/* Largest k accepted by the streaming top-k kernels (k+1 <= 128 entries kept
 * per row in registers); larger k is served by scedar_col_sort. */
int scedar_max_k(void);

/* Number of kernels this library has launched in this process (monotonic). */
long long scedar_launch_count(void);

/* ------------------------------------------------------------------------
 * Prepared cells: the device-resident working copy of x that the Gram kernels
 * read (rows padded to a multiple of 128 cells, genes to a multiple of 16,
 * zero filled; centred on the row mean for correlation, sdm.py:1286) plus the
 * per-cell statistic of the metric epilogue (sqrt(1/|x|^2) with 0 for zero
 * vectors, sdm.py:1249-1261; |x|^2 for euclidean), taken from the diagonal of
 * the same Gram contraction as the reference does (np.diag(pdot_prod)).
 * Replaces the input contract of SampleFeatureMatrix.__init__, sfm.py:43-75.
 * ---------------------------------------------------------------------- */
typedef struct scedar_cells scedar_cells_t;

/* x: [n, p] float64, row stride ldx (elements). */
int scedar_cells_from_dense(const double* x, int64_t n, int64_t p, int64_t ldx,
                            int metric, void* stream, scedar_cells_t** out);

/* CSR cells (cosine / euclidean only, as sdm.py:1196-1204): indptr has n+1
 * entries of int32 (indptr_is_64 = 0) or int64 (1); duplicates are summed. */
int scedar_cells_from_csr(const void* indptr, int indptr_is_64,
                          const int32_t* indices, const double* data,
                          int64_t n, int64_t p, int metric, void* stream,
                          scedar_cells_t** out);

/* Incremental form, for cells that arrive in pieces (e.g. while the host ->
 * device copy of x is still in flight): scedar_cells_alloc reserves the
 * working copy, scedar_cells_fill_rows prepares the cells [row_begin, row_end)
 * from x_rows (pointer to cell row_begin, row stride ldx) exactly as
 * scedar_cells_from_dense does.  Ranges are 128-aligned (the last one ends at
 * n) and may be filled in any order; a cell must be filled before a kernel
 * reads it. */
int scedar_cells_alloc(int64_t n, int64_t p, int metric, void* stream,
                       scedar_cells_t** out);
int scedar_cells_fill_rows(scedar_cells_t* cells, const double* x_rows,
                           int64_t ldx, int64_t row_begin, int64_t row_end,
                           void* stream);
/* The same with the Gram diagonal -- the per-cell statistic the DMMA and 3xFP16
 * kernels read -- optional (with_stat = 0): many small ranges then cost one
 * preparation pass each instead of a preparation pass and a 128-row tile
 * contraction.  The INT8 split kernels take their statistic from the byte
 * slices and need nothing else; before any other kernel uses the handle, run
 * scedar_cells_stat_rows over the rows that were filled without it. */
int scedar_cells_fill_rows_ex(scedar_cells_t* cells, const double* x_rows,
                              int64_t ldx, int64_t row_begin, int64_t row_end,
                              int with_stat, void* stream);
int scedar_cells_stat_rows(scedar_cells_t* cells, int64_t row_begin,
                           int64_t row_end, void* stream);

/* Stream-ordered on the stream the cells were created on (work that reads the
 * cells on another stream must have been synchronised with it). */
int scedar_cells_free(scedar_cells_t* cells);
int64_t scedar_cells_n(const scedar_cells_t* cells);
int64_t scedar_cells_p(const scedar_cells_t* cells);
int scedar_cells_metric(const scedar_cells_t* cells);

/* ------------------------------------------------------------------------
 * Distance matrix: rows [row_begin, row_end) of the n x n matrix `_d`
 * (sdm.py:1193-1221) into d_out[(row_end-row_begin), n] with row stride ldd.
 * The epilogue applies the metric formula in the reference's operation order
 * and num_correct_dist_mat's clean-up (sdm.py:215-221): negatives -> 0, exact
 * zero diagonal, value of the lower triangle on both sides.
 * ---------------------------------------------------------------------- */
int scedar_pdist_stripe(const scedar_cells_t* cells, int precision,
                        int64_t row_begin, int64_t row_end, double* d_out,
                        int64_t ldd, void* stream);

/* Multi-GPU form over peer memory: rank r of `world` (one process per GPU)
 * owns rows [stripe_begin[r], stripe_begin[r+1]) (128-aligned, covering [0,n))
 * in stripe_ptr[r], all [rows_r, n] with row stride ldd and all mapped on THIS
 * device (own stripe + peers opened with scedar_stripe_open).  Every
 * off-diagonal 128 x 128 tile pair of the whole matrix is computed once -- by
 * the owner of the higher row tile when the tile indices sum to an even number,
 * the lower otherwise -- and the epilogue stores the tile into the local stripe
 * and its transpose straight into the other owner's stripe over NVLink, so N
 * GPUs do n(n+1)p flops in total instead of 2n^2p.  stripe_begin / stripe_ptr
 * are HOST arrays.  The caller must barrier all ranks on their streams before
 * anyone reads a stripe and before the next call writes it. */
int scedar_pdist_stripe_p2p(const scedar_cells_t* cells, int precision, int rank,
                            int world, const int64_t* stripe_begin,
                            double* const* stripe_ptr, int64_t ldd,
                            void* stream);

/* Row-range form of scedar_pdist_stripe_p2p for cells that arrive piece by piece
 * (scedar_cells_alloc + scedar_cells_fill_rows on every rank as the pieces of x
 * are uploaded and gathered): the lower-triangle tiles of the rows [row_begin,
 * row_end) -- which need the cells [0, row_end) only -- dealt column block by
 * column block over the ranks, each stored with its transpose into the owners'
 * stripes.  Called for consecutive ranges in ascending order on every rank it
 * yields the matrix of scedar_pdist_stripe_p2p bit for bit.  reserve_sms SMs
 * are left idle for the collective that gathers the next piece.  INT8 split
 * kernels only (SCEDAR_PRECISION_FP64_I8 / _FP64_I8S6). */
int scedar_pdist_p2p_rows(const scedar_cells_t* cells, int precision, int rank,
                          int world, const int64_t* stripe_begin,
                          double* const* stripe_ptr, int64_t ldd,
                          int64_t row_begin, int64_t row_end, int reserve_sms,
                          void* stream);

/* Piece form of scedar_pdist_stripe_p2p, for x that sits row-sharded BY STRIPE in
 * the ranks' host memories and is uploaded and exchanged in n_pieces pieces, piece
 * q being the q-th part of EVERY stripe (piece_of_tile[T], T < ceil(n / 128): the
 * piece that brings row tile T; scedar_cells_alloc + scedar_cells_fill_rows as
 * the rows land).  Launch `piece` contracts the pairs of row tiles whose later
 * tile arrives with that piece, owned and oriented as in scedar_pdist_stripe_p2p
 * (tile into the rank's own stripe, mirrored tile over NVLink), so every rank
 * gets the same share of every launch while the next piece is on its way.
 * Launches 0 .. n_pieces-1 on every rank yield the matrix of
 * scedar_pdist_stripe_p2p bit for bit.  INT8 split kernels only. */
int scedar_pdist_p2p_pieces(const scedar_cells_t* cells, int precision, int rank,
                            int world, const int64_t* stripe_begin,
                            double* const* stripe_ptr, int64_t ldd,
                            const int32_t* piece_of_tile, int piece, int n_pieces,
                            int reserve_sms, void* stream);

/* Stripe buffers that can be mapped by the peer processes of one node
 * (cudaMalloc + CUDA IPC).  handle64 is a 64-byte opaque blob to ship to the
 * peers (e.g. torch.distributed.all_gather_object). */
int scedar_stripe_alloc(size_t bytes, void** dev_ptr);
int scedar_stripe_free(void* dev_ptr);
int scedar_stripe_export(void* dev_ptr, void* handle64);
int scedar_stripe_open(const void* handle64, void** peer_ptr);
int scedar_stripe_close(void* peer_ptr);

/* The rows [row_begin, row_end) of the LOWER triangle (columns <= row) of the
 * whole matrix d [n, n] (FP64 path), each tile written with its transpose.
 * It reads only the cells [0, row_end): called for consecutive row ranges in
 * ascending order it yields the matrix of scedar_pdist_symmetric bit for bit
 * while later cells are still being uploaded (scedar_cells_fill_rows). */
int scedar_pdist_lower_rows(const scedar_cells_t* cells, int64_t row_begin,
                            int64_t row_end, double* d, int64_t ldd,
                            void* stream);

/* scedar_pdist_lower_rows with the Gram kernel chosen by `precision`
 * (SCEDAR_PRECISION_FP64, _FP64_I8 or _FP64_I8S6). */
int scedar_pdist_lower_rows_prec(const scedar_cells_t* cells, int precision,
                                 int64_t row_begin, int64_t row_end, double* d,
                                 int64_t ldd, void* stream);

/* Whole matrix at once, computing only the lower-triangle tiles and writing
 * each tile and its transpose (half the contraction work). */
int scedar_pdist_symmetric(const scedar_cells_t* cells, int precision,
                           double* d_out, int64_t ldd, void* stream);

/* ------------------------------------------------------------------------
 * PCA / truncated-SVD producer of the kNN path (scedar/eda/sdm.py:1313-1334,
 * sklearn PCA(svd_solver="auto", random_state=17) / TruncatedSVD): the n^2- and
 * n p^2-sized pieces.  The p x p (or n x n) symmetric eigenproblem in between
 * is the caller's (cuSOLVER through torch in scedar_b200/pca.py).
 * ---------------------------------------------------------------------- */
/* mean of every column over the real rows of the working copy: mean_out [p] */
int scedar_cells_col_means(const scedar_cells_t* cells, double* mean_out,
                           void* stream);
/* working copy -= shift[column] (column centring), in place */
int scedar_cells_shift_cols(scedar_cells_t* cells, const double* shift,
                            void* stream);
/* new handle holding the transposed working copy (rows = features, columns =
 * samples), optionally minus shift[feature]; metric as given (no statistic) */
int scedar_cells_transposed(const scedar_cells_t* cells, const double* shift,
                            void* stream, scedar_cells_t** out);
/* bare Gram matrix of the prepared rows, out [n, n] (for transposed cells: the
 * scatter matrix XtX), FP64 tensor-core kernel, lower tiles mirrored */
int scedar_gram_raw(const scedar_cells_t* cells, double* out, int64_t ldd,
                    void* stream);
/* bare cross products A.Bt of two handles with the same column count:
 * out [n_a, n_b] */
int scedar_gram_raw_cross(const scedar_cells_t* a, const scedar_cells_t* b,
                          double* out, int64_t ldd, void* stream);

/* ------------------------------------------------------------------------
 * k nearest neighbours of rows [row_begin, row_end) without materialising D:
 * Gram tile -> metric epilogue -> streaming top-(k+1) per row, ordered by
 * (distance, index) -- the order of np.argsort(kind="mergesort")
 * (sdm.py:1298-1305).  idx_out / dist_out: [(row_end-row_begin), k].
 * Replaces s_knn_ind_lut (sdm.py:751-790) and _s_knns_skl (sdm.py:1053-1083).
 * ---------------------------------------------------------------------- */
int scedar_knn_stripe(const scedar_cells_t* cells, int precision, int k,
                      int exclude, int64_t row_begin, int64_t row_end,
                      int64_t* idx_out, double* dist_out, void* stream);

/* Same selection from an already materialised stripe of D:
 * d is [(row_end-row_begin), n_cols] with row stride ldd; row r of d is cell
 * row_begin + r.  (NearestNeighbors(metric="precomputed"), sdm.py:1070-1072.) */
int scedar_knn_from_d(const double* d, int64_t ldd, int64_t n_cols,
                      int64_t row_begin, int64_t row_end, int k, int exclude,
                      int64_t* idx_out, double* dist_out, void* stream);

/* Block minima, an optional by-product of the INT8 split launches
 * (SCEDAR_PRECISION_FP64_I8 / _I8S6): while a tile of D is in registers the
 * kernel also stores the minimum of every aligned run of 32 entries of each
 * row, and the selection below then reads n/32 minima plus the ~k+1 runs that
 * can hold a neighbour instead of the whole row (same lists, bit for bit).
 * stripe_bmin[r] is the buffer of output stripe r -- the stripes of the launch
 * that follows: one for the single-device forms, `world` for the peer-store
 * forms -- laid out [rows of the stripe][ldb], ldb >= ceil(n / 32), row 0 =
 * first row of the stripe.  world = 0 detaches the buffers. */
int scedar_cells_set_blockmin(scedar_cells_t* cells, double* const* stripe_bmin,
                              int world, int64_t ldb);
/* scedar_knn_from_d over rows that an INT8 split launch on `cells` has written
 * together with their block minima (an error if none has since
 * scedar_cells_set_blockmin).  bmin: [(row_end-row_begin), ldb]. */
int scedar_knn_from_d_blockmin(const scedar_cells_t* cells, const double* d,
                               int64_t ldd, const double* bmin, int64_t ldb,
                               int64_t row_begin, int64_t row_end, int k,
                               int exclude, int64_t* idx_out, double* dist_out,
                               void* stream);

/* ------------------------------------------------------------------------
 * Full stable sort of every column of a symmetric D (sdm.py:1289-1305,
 * np.sort / np.argsort(kind="mergesort", axis=0)).  Column j of a symmetric D
 * is row j, so the sort works on rows: d is a stripe [rows, n_cols] (row
 * stride ldd) and sorted_out / argsorted_out are [rows, n_cols], entry
 * (j, r) = r-th nearest of the j-th cell of the stripe -- the TRANSPOSE of the
 * reference's (rank, cell) arrays.  Either output may be NULL.
 * ---------------------------------------------------------------------- */
int scedar_col_sort(const double* d, int64_t ldd, int64_t rows, int64_t n_cols,
                    double* sorted_out, int64_t* argsorted_out, void* stream);

/* ------------------------------------------------------------------------
 * num_correct_dist_mat (sdm.py:187-222) in place on a user-supplied n x n
 * matrix: *flags_out (device int32) receives SCEDAR_NCDM_* bits for the two
 * warnings; then negatives -> 0, diagonal -> 0, optional clamp to upper_bound,
 * upper triangle := lower triangle.
 * ---------------------------------------------------------------------- */
int scedar_num_correct_dist_mat(double* d, int64_t ldd, int64_t n,
                                int has_upper_bound, double upper_bound,
                                int32_t* flags_out, void* stream);

/* ------------------------------------------------------------------------
 * Consumers directly behind the path: everything below works
 * on device buffers so the kNN lists / D never round-trip through the host.
 * ---------------------------------------------------------------------- */

/* kNN lists -> CSR connectivity matrix (s_knn_connectivity_matrix,
 * sdm.py:934-947: coo_matrix((dist, (source, target))).tocsr()).  idx / dist
 * are [n, k] (cell i's neighbours in row i); indptr_out [n+1] = i*k,
 * indices_out / data_out [n*k] with the columns of a row ascending (what
 * scipy's tocsr() + sum_duplicates() yields) and exact-zero distances stored
 * as -inf (sdm.py:942).  k <= 128. */
int scedar_knn_connectivity_csr(const int64_t* idx, const double* dist,
                                int64_t n, int k, int64_t* indptr_out,
                                int32_t* indices_out, double* data_out,
                                void* stream);

/* CSR data -> affinity edge weights of knn_conn_mat_to_aff_graph
 * (sdm.py:1085-1093): -inf -> 0, then (max(w) - w) * aff_scale. */
int scedar_knn_affinity(const double* data, int64_t nnz, double aff_scale,
                        double* weights_out, void* stream);

/* RareSampleDetection on a resident D (scedar/knn/detection.py:81-123).
 * masked_kth: for rows [row_begin, row_end) of the stripe d (row r = cell
 * row_begin + r; row stride ldd) that are alive, the value at 0-based `rank`
 * of the row restricted to the alive columns -- np.sort(curr_dist_mat,
 * axis=0)[i_k, s] of the shrinking sub-matrix (detection.py:100, 111, 119)
 * without slicing or sorting; +inf for dead rows.  rank < 128.
 * alive: uint8 [n_cols]. */
int scedar_masked_kth(const double* d, int64_t ldd, int64_t n_cols,
                      int64_t row_begin, int64_t row_end, const uint8_t* alive,
                      int rank, double* kth_out, void* stream);
/* One detection iteration (detection.py:63-66, 111-115): alive cells whose
 * kth[i*ldk] <= cutoff stay alive and get last_kept[i] = iter, the others
 * die.  *n_kept_out (device uint64) is incremented by the number kept. */
int scedar_rare_update(const double* kth, int64_t ldk, int64_t n, double cutoff,
                       int iter, uint8_t* alive, int32_t* last_kept,
                       unsigned long long* n_kept_out, void* stream);
/* max over the alive entries of a non-negative strided vector (d_max,
 * detection.py:45, 101); alive may be NULL.  *max_out: device double,
 * initialised to 0 by the caller. */
int scedar_masked_max(const double* v, int64_t ldv, int64_t n,
                      const uint8_t* alive, double* max_out, void* stream);
/* ascending indices of the alive cells -> sel_out (capacity >= #alive) */
int scedar_compact_alive(const uint8_t* alive, int64_t n, int64_t* sel_out,
                         void* stream);
/* out[r, :] = x[sel[r], :] (detection.py:71: curr_sdm._x[kept, :]) */
int scedar_gather_rows(const double* x, int64_t ldx, int64_t p,
                       const int64_t* sel, int64_t m, double* out, int64_t ldo,
                       void* stream);

/* One iteration of FeatureImputation._impute_features_runner
 * (scedar/knn/imputation.py:56-101) with statistic_fun = np.median: for every
 * cell s and gene f with x_curr[s,f] < min_present_val whose k neighbours
 * knn[s, :] hold >= n_do_i values >= min_present_val, x_next[s,f] = median of
 * the k neighbour values and idc[s,f] = iter; every other entry is copied.
 * Reads x_curr only (the reference edits a copy).  stats_out (device
 * uint64[2], may be NULL) receives the number of entries and of cells with
 * idc > 0 after the iteration (imputation.py:105-109).  k <= 200.
 * Only the cells [row_begin, row_end) are processed (a rank's row stripe; the
 * whole matrix is 0, n): x_curr, x_next, knn and idc are the full [n, .]
 * buffers, rows outside the range are neither written nor counted. */
int scedar_impute_iter(const double* x_curr, int64_t ldx, int64_t n, int64_t p,
                       const int64_t* knn, int k, int n_do_i,
                       double min_present_val, int iter, int64_t row_begin,
                       int64_t row_end, double* x_next, int64_t ldn,
                       int32_t* idc, int64_t ldi,
                       unsigned long long* stats_out, void* stream);

/* ------------------------------------------------------------------------
 * Clustering consumers of D: the data movement between
 * the distance matrix and the sequential CPU linkage, on the device.
 * ---------------------------------------------------------------------- */

/* Condensed form of a symmetric D: scipy.spatial.distance.squareform(dmat) as
 * HClustTree.hclust_linkage calls it before sch.linkage (sdm.py:1666-1668).
 * d is the stripe [(row_end-row_begin), n] of D (row r = cell row_begin + r,
 * row stride ldd); the strict upper triangle of those rows, row after row, is
 * the contiguous segment [base(row_begin), base(row_end)) of the condensed
 * vector, base(i) = i(2n-i-1)/2; condensed_out points at entry
 * base(row_begin) (the whole vector for row_begin = 0).  Symmetry of D is the
 * caller's contract (the Gram kernels and scedar_num_correct_dist_mat both
 * establish it bit for bit); what scipy's validity check can still reject is
 * OR-ed into *flags_out (device int32, zeroed by the caller, shared by the
 * calls of all stripes): SCEDAR_SQF_*. */
int scedar_squareform_rows(const double* d, int64_t ldd, int64_t n,
                           int64_t row_begin, int64_t row_end,
                           double* condensed_out, int32_t* flags_out,
                           void* stream);

/* Sub-matrix out[r, c] = d[row_sel[r], col_sel[c]] (m x c, row stride ldo) of
 * a matrix / stripe d [n_rows, n_cols] (row stride ldd): the slicing
 * self._d[s][:, s] of SampleDistanceMatrix.ind_x (sdm.py:684-692) and
 * self._sdm._d[a][:, b] of MIRAC (scedar/cluster/mirac.py:316-331).  Indices
 * are int64, already non-negative (NumPy's wrap-around is the caller's);
 * an index outside [0, n_rows) / [0, n_cols) sets SCEDAR_GATHER_BAD_INDEX in
 * *flags_out (device int32, zeroed by the caller) and a NaN in its slot. */
int scedar_d_gather_block(const double* d, int64_t ldd, int64_t n_rows,
                          int64_t n_cols, const int64_t* row_sel, int64_t m,
                          const int64_t* col_sel, int64_t c, double* out,
                          int64_t ldo, int32_t* flags_out, void* stream);

/* ------------------------------------------------------------------------
 * One-shot forms (prepare + run + free), the signatures SURVEY.md 8b lists.
 * ---------------------------------------------------------------------- */
int scedar_pdist_dense(const double* x, int64_t n, int64_t p, int metric,
                       int precision, int64_t row_begin, int64_t row_end,
                       double* d_out, void* stream);
int scedar_pdist_csr(const void* indptr, int indptr_is_64,
                     const int32_t* indices, const double* data, int64_t n,
                     int64_t p, int metric, int precision, int64_t row_begin,
                     int64_t row_end, double* d_out, void* stream);
int scedar_knn_dense(const double* x, int64_t n, int64_t p, int metric,
                     int precision, int k, int exclude, int64_t row_begin,
                     int64_t row_end, int64_t* idx_out, double* dist_out,
                     void* stream);

/* ------------------------------------------------------------------------
 * Host-buffer forms: x, d_out, idx_out, dist_out are HOST pointers; the call
 * does the host<->device copies itself on `device` and returns when the
 * outputs are complete.  D is produced and copied back stripe by stripe, so
 * the n x n matrix never has to fit in device memory.
 * ---------------------------------------------------------------------- */
int scedar_pdist_dense_host(const double* x, int64_t n, int64_t p, int metric,
                            int precision, double* d_out, int device);
int scedar_knn_dense_host(const double* x, int64_t n, int64_t p, int metric,
                          int precision, int k, int exclude, int64_t* idx_out,
                          double* dist_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* SCEDAR_B200_H */
