/*
 * dfol_b200.h -- C ABI of libdfol_b200.so: the B200 (sm_100a) kernels of the differentiable-FOL reasoning path.
 *
 * The reference (microsoft/DFOL-VQA) is pure Python/PyTorch and has no FFI; each entry point below states the
 * reference interface it replaces (file:line under src/ of the reference).  INTEGRATION.md shows the binding a
 * reference maintainer would add (ctypes).
 *
 * Conventions:
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *  - the caller owns all buffers (including workspaces); the library never allocates, frees or keeps a pointer;
 *  - every call is asynchronous on the given stream (cudaStream_t passed as void*); no host synchronisation;
 *  - return value 0 = success, >0 = cudaError_t of the launch, <0 = argument error; dfol_last_error() returns a
 *    thread-local message for the last non-zero return;
 *  - fp32 tables; indices int32 (int64 for element offsets of the big tables).
 *
 * Device layout of a scene (one program batch = B questions = B images, image b has N_b objects):
 *  - object rows are concatenated raggedly: image b owns rows [obj_row[b], obj_row[b+1]);
 *  - pair rows enumerate (b, s, o) for ALL s,o in [0,N_b) (self pairs included; their table entries are -30):
 *    image b owns pair rows [pair_row[b], pair_row[b+1]), local index l = s*N_b + o;
 *  - attribute table: image block b starts at element attr_blk[b] and is [C][attr_stride[b]] (concept-major,
 *    attr_stride[b] = N_b rounded up to 4) -- a Filter reads one contiguous row of N_b floats;
 *  - relation table: image block b starts at element rel_blk[b] and is [nR][rel_stride[b]] with
 *    rel_stride[b] = N_b*N_b rounded up to 4 -- a Relate reads one contiguous N_b x N_b tile.
 */
#ifndef DFOL_B200_H
#define DFOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFOL_ABI_VERSION 6

/* activation codes (RegularMLP / EmbeddingLayer, gqa_interpreter_experiments.py:28-33, :71-72) */
#define DFOL_ACT_NONE 0
#define DFOL_ACT_ELU 1
#define DFOL_ACT_SIGMOID 2
#define DFOL_ACT_LOGSIGMOID 3

/* epilogue multiplier codes for backward GEMMs: dZ = dH * act'(.) evaluated from the saved activation */
#define DFOL_MUL_NONE 0
#define DFOL_MUL_SIGMOID_GRAD 1 /* h*(1-h) */
#define DFOL_MUL_ELU_GRAD 2     /* h>0 ? 1 : h+1 */

int dfol_version(void);
const char* dfol_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Dense contractions of the visual oracle.
 * Replaces the nn.Linear + activation chains of RegularMLP / EmbeddingLayer (gqa_interpreter_experiments.py:
 * 18-77) as called from featurize_scene (nsvqa/data/batch_gqa_boxfeatures_pipeline.py:204) and
 * ClassifierOracle.compute_all_log_likelihood_2 (nsvqa/nn/vision/classifier_oracle.py:145-156), and their
 * autograd backward.
 *
 * C = epilogue( sum_k A(m,k) * B(k,n) ),  A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]  (fp32).
 * epilogue: + bias[n]; activation `act`; * multiplier from mul_src[m*ld_mul + n]; then
 *   store == 0: C[m*ldc + n]  (accumulate != 0 adds to the existing value)
 *   store == 1: per-image transposed table store, C[img_blk[b] + n*img_stride[b] + (m - img_row[b])] with
 *               b = row_img[m]; if img_n != NULL rows with s == o (self pairs) are written as diag_value.
 * split_k > 1 splits K over grid.z and atomically adds into C (C must be zeroed; bias/act/mul must be off).
 * fp32 SIMT kernel: the parity path (bit-level behaviour of fp32 FMA accumulation).
 */
int dfol_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                  int64_t ldc, const float* bias, int M, int N, int K, int act, int accumulate, int split_k,
                  const float* mul_src, int64_t ld_mul, int mul_mode, int store, const int32_t* row_img,
                  const int32_t* img_row, const int64_t* img_blk, const int32_t* img_stride, const int32_t* img_n,
                  float diag_value, void* stream);

/* bf16 tensor-core GEMM (tcgen05.mma, TMA-staged operands, fp32 accumulation in TMEM), same epilogue contract.
 * A is [M,K] row-major bf16 (lda elements), B is [N,K] row-major bf16 (ldb elements) i.e. C = A * B^T.
 * K must be a multiple of 64 (pad with zeros), lda/ldb >= K and multiples of 8, A/B 16-byte aligned.
 * store == 0: row-major C with ldc elements per row, bf16 if out_bf16 != 0 else fp32; columns N <= n < ldc are
 * written as ZERO so that C can be the (K-padded) A operand of the next layer.  store == 1: fp32 table store.
 * The TMA tensor maps are encoded inside the call (host side, no device work). */
int dfol_gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                      const float* bias, int M, int N, int K, int act, int out_bf16, int store,
                      const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                      const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream);

/* Backward contractions of the tensor-core mode (bf16 operands, fp32 accumulation):
 * dgrad: dX[M,N] (bf16, ld lddx) = (dZ[M,K] . Wt[N,K]^T) * act'(h_saved[m,n])  -- Wt is the TRANSPOSED weight,
 *        mul_mode = DFOL_MUL_* evaluated from the saved bf16 activation output (epilogue multiplier); columns
 *        N <= n < store_cols (0 = lddx) are written as zero, columns beyond store_cols are left untouched (dX may
 *        be a column block of a wider concatenated operand);
 * wgrad: C[M,N] (fp32) += A[K,M]^T . B[K,N] with the reduction over ROWS (K = pair / object rows): MN-major UMMA
 *        operands straight from the row-major activations, split-K over CTAs, red.global.add.f32 epilogue. */
int dfol_gemm_bf16_tc_dgrad(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX, int64_t lddx,
                            int store_cols, int M, int N, int K, const void* h_saved, int64_t ldh, int mul_mode,
                            float keep, void* stream);
/* keep (here, in dfol_pair_layer_dgrad_cluster and in dfol_table_layer_bwd_tc) = 1 - dropout p of the layer whose saved
 * activation is passed: 1 = no dropout; < 1 = the saved tensor holds the activation AFTER dropout (0 where dropped,
 * h / keep where kept), so act' is taken at saved * keep and the gradient carries the mask factor (saved != 0) / keep. */
int dfol_gemm_bf16_tc_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                            int64_t K, void* stream);
/* same contraction with the M rows of the result split into up to three segments of seg_rows rows (multiple of 128)
 * that accumulate into different gradient tensors: one launch for the weight gradients of all first layers that read
 * the object features (A = [dZ1_attr | dU | dV], B = obj). */
int dfol_gemm_bf16_tc_wgrad_seg(const void* A, int64_t lda, const void* B, int64_t ldb, float* C0, int64_t ldc0,
                                float* C1, int64_t ldc1, float* C2, int64_t ldc2, int seg_rows, int M, int N, int64_t K,
                                void* stream);

/* fp32 parity mode on the tensor cores (split-bf16: x = h + m + l, six product terms, one bf16 GEMM over operands
 * concatenated along K, fp32 TMEM accumulation; replaces the fp32 nn.Linear arithmetic of classifier_oracle.py:145-156
 * with results at fp32 level).  dfol_split3_bf16 writes the concatenated operand of an fp32 matrix: pattern 0 (A side:
 * h l m h m h) or 1 (B side: l h m m h h; smallest product terms first, hh last); stacked == 0: dst[r, k*Kp + c] (K-concatenated, zero padded to Kp, for
 * dfol_gemm_bf16_tc_exact); stacked != 0: dst[(k*rows + r), c] (row-concatenated, for dfol_gemm_bf16_tc_wgrad).
 * dfol_gemm_bf16_tc_exact = dfol_gemm_bf16_tc with fp32 output, accurate (expf / log1pf) activations and an explicit
 * stored width. */
int dfol_split3_bf16(const float* src, int64_t lds, int64_t rows, int cols, void* dst, int64_t ldd, int Kp, int pattern,
                     int stacked, void* stream);
int dfol_gemm_bf16_tc_exact(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                            int store_cols, const float* bias, int M, int N, int K, int act, int store,
                            const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                            const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream);

/* Batched operand preparation: job j casts (and, if transpose != 0, transposes) the fp32 view src[rows][cols]
 * (row stride lds) into columns [0, dcols) of the bf16 rows dst[out_rows][ldd], zero beyond the source extent.
 * `jobs` is a DEVICE array of job_num records of dfol_cast_job_size() bytes:
 *   { const float* src; int64 lds; int32 rows, cols; bf16* dst; int64 ldd; int32 out_rows, transpose, dcols, pad }.
 * max_elements = largest out_rows*dcols of the jobs (grid sizing).  One launch refreshes every weight operand. */
int dfol_cast_jobs(const void* jobs, int job_num, int64_t max_elements, void* stream);
int dfol_cast_job_size(void);

/* Persistent, weights-resident tensor-core GEMM for the pair-level layers of the relation network
 * (RegularMLP over the (P, .) pair matrix, gqa_interpreter_experiments.py:167; classifier_oracle.py:150-154).
 * One CTA per SM keeps the whole weight matrix B[N,K] (N <= 320, K <= 320) in shared memory and streams the 128-row
 * tiles of A through a TMA ring; accumulators live in TMEM (double buffered when N <= 256).
 * dfol_pair_layer_fwd_tc: C = act(A.B^T + bias) stored as bf16 (C == NULL: not stored), columns N <= n < store_cols
 *   zero.  If slot_wrow != NULL (act must be sigmoid) the epilogue also evaluates the demand-driven relation
 *   columns of dfol_rel_slots_fwd from the fp32 activations while they are in registers:
 *   ll[slot_blk[b] + k*rel_stride[b] + l] = logsigmoid(C[m,:] . W_emb[wrow,:] + b_emb[wrow]), b = row_img[m],
 *   l = m - img_row[b], self pairs = diag_value.
 * dfol_pair_layer_dgrad_tc: same contract as dfol_gemm_bf16_tc_dgrad. */
int dfol_pair_layer_fwd_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int store_cols,
                           const float* bias, int M, int N, int K, int act, const float* W_emb, int64_t ldw,
                           const float* b_emb, const int32_t* slot_wrow, const int32_t* img_slot, int max_slots,
                           const int64_t* slot_blk, const int32_t* rel_stride, const int32_t* row_img,
                           const int32_t* img_row, const int32_t* img_n, float diag_value, float* ll, void* stream);
int dfol_pair_layer_dgrad_tc(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX, int64_t lddx,
                             int store_cols, int M, int N, int K, const void* h_saved, int64_t ldh, int mul_mode,
                             void* stream);

/* Cluster-of-two variant of the pair-layer GEMMs (same arithmetic contract, bf16 row-major C, N <= 384): two CTAs on
 * two SMs split the output columns, each keeps half of B resident; every A block is fetched once per cluster
 * (TMA .multicast::cluster), accumulators are double buffered in TMEM, and the epilogue moves its operand
 * (h_saved of the dgrad) and its result through 128B-swizzled shared tiles with TMA loads / stores, so global
 * memory only ever sees full-line bulk transfers.  store_cols (0 = ldc) columns are written, zero beyond N. */
int dfol_pair_layer_fwd_cluster(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                int store_cols, const float* bias, int M, int N, int K, int act, void* stream);
int dfol_pair_layer_dgrad_cluster(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX, int64_t lddx,
                                  int store_cols, int M, int N, int K, const void* h_saved, int64_t ldh, int mul_mode,
                                  float keep, void* stream);
/* The same dgrad with the WEIGHT GRADIENT of the layer fused in (backward of nn.Linear, classifier_oracle.py:149-154 via
 * autograd): dW[k, n] += sum_rows dZ[row, k] * h_saved[row, n] for k < k_real (fp32, red.global.add), accumulated in a
 * second TMEM accumulator from the operands the dgrad already holds in shared memory (MN-major tcgen05 MMAs over the
 * tile's 128 rows) -- dZ and h_saved are read from HBM once for both products.  Needs 65..256 output columns (N) and
 * K <= 384 (128 + K TMEM columns). */
int dfol_pair_layer_dgrad_wgrad_cluster(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX,
                                        int64_t lddx, int store_cols, int M, int N, int K, const void* h_saved,
                                        int64_t ldh, int mul_mode, float keep, float* dW, int64_t lddw, int k_real,
                                        void* stream);

/* fp32 -> bf16 cast with row padding: dst[r*ldd + c] = bf16(src[r*lds + c]) for c < cols, 0 for cols <= c < ldd */
int dfol_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Featurizer tail + pairwise hidden layer.
 * dfol_box_position: obj[t, F:F+4] = [x,y,w,h] / max([W,H,W,H],1)   (batch_gqa_boxfeatures_pipeline.py:208-211)
 * dfol_pair_hidden_fwd: first layer of the relation network on the pair feature
 *   [obj_s | obj_o | dist | asin(dy/dist) | sign(x_o-x_s) | sign(y_o-y_s)]  (:257-279), evaluated WITHOUT
 *   materialising the (P,2F+12) pair matrix:  h = act( U[s] + V[o] + Wg . geo(s,o) + b )  where
 *   U = obj.Ws^T, V = obj.Wo^T are the two halves of the first Linear (uv[t, 0:H] = U, uv[t, H:2H] = V) and Wg its
 *   last four input columns (wg[h*ldw + j], j<4).  One output row per pair row (self pairs included); fp32, or
 *   bf16 with columns H..ldh zero-filled when out_bf16 != 0 (operand of the next tensor-core GEMM).
 * dfol_pair_hidden_bwd: given dH (pair rows x H) and the saved H: dZ = dH*act'(H); duv[s,0:H] += sum_o dZ,
 *   duv[o,H:2H] += sum_s dZ, dwg[h*ldw + j] += sum dZ*geo_j, db[h] += sum dZ  (duv, dwg, db must be zeroed).
 */
int dfol_box_position(const float* features, int64_t ldf, int feature_dim, float* obj, int64_t ldo, int out_col,
                      int64_t rows, void* stream);
int dfol_pair_hidden_fwd(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos, const float* wg,
                         int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H, int act, int out_bf16,
                         const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n, int image_num,
                         int max_n, void* stream);
int dfol_pair_hidden_bwd(const float* dh, int64_t lddh, const float* h_saved, int64_t ldh, const float* obj_pos,
                         int64_t ldpos, float* duv, int64_t lduv, float* dwg, int64_t ldw, float* dbias, int H,
                         int act, const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n,
                         int image_num, void* stream);

/* Tensor-core mode (bf16 activations), single-pass HBM-bound kernels:
 * dfol_obj_finish: writes the box position columns obj[t, F:F+4] (as dfol_box_position) and the bf16 operand copy
 *   obj16[t, 0:ld16) = bf16(obj[t, :]) with zero K-padding.
 * dfol_pair_hidden_fwd_tc: h = elu(U[s]+V[o]+Wg.geo+b) in bf16 (columns H..ldh zero) and, if geo_out != NULL, the
 *   pair geometry table geo_out[pair] = float4(dist, asin, sign x, sign y) that the backward kernel re-uses.
 * dfol_pair_hidden_bwd_tc: from dz (bf16, activation derivative already applied by the dgrad epilogue) and geo:
 *   du_out[t, 0:H] = bf16(sum_o dz), dv_out[t, 0:H] = bf16(sum_s dz) (row stride ldo elements; WRITTEN),
 *   dwg[h*ldw + k] += sum dz*geo_k, dbias[h] += sum dz (atomics; zeroed by the caller). */
int dfol_obj_finish(const float* features, int64_t ldf, int feature_dim, float* obj, int64_t ldo, int F, void* obj16,
                    int64_t ld16, int64_t rows, void* stream);
int dfol_pair_hidden_fwd_tc(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos, const float* wg,
                            int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H, void* geo_out,
                            const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n, int image_num,
                            int max_n, void* stream);
int dfol_pair_hidden_bwd_tc(const void* dz, int64_t lddz, const void* geo, void* du_out, void* dv_out, int64_t ldo,
                            float* dwg, int64_t ldw, float* dbias, int H, const int32_t* pair_row,
                            const int32_t* obj_row, const int32_t* img_n, int image_num, int max_n, void* stream);

/* bf16 variant: dz already carries the activation derivative (dgrad epilogue); duv is WRITTEN (no memset needed). */
int dfol_pair_hidden_bwd_bf16(const void* dz, int64_t lddz, const float* obj_pos, int64_t ldpos, float* duv,
                              int64_t lduv, float* dwg, int64_t ldw, float* dbias, int H, const int32_t* pair_row,
                              const int32_t* obj_row, const int32_t* img_n, int image_num, int max_n, void* stream);
int dfol_colsum_bf16(const void* X, int64_t ldx, int64_t M, int N, float* out, void* stream);

/* in place dH[m,n] *= act'(.) evaluated from the saved activation output H[m,n] (act = DFOL_ACT_*) */
int dfol_act_grad_mul(float* dH, int64_t lddh, const float* H, int64_t ldh, int64_t rows, int cols, int act,
                      void* stream);

/* column sums: out[n] += sum_m X[m*ldx + n]  (bias gradients) */
int dfol_colsum(const float* X, int64_t ldx, int64_t M, int N, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Program interpreter: executes compiled FOL programs, one thread block per question.
 * Replaces BatchInterpreterBase.forward's execution loop (nsvqa/nn/interpreter/batch_base_interpreter.py:
 * 147-172), the op modules (batch_gqa_ops.py:160-780, batch_base_ops.py:42-215, 311-405, 483-596), the
 * per-op oracle gather (nsvqa/nn/vision/classifier_oracle.py:44-137), BatchVariableSet.log_probability /
 * gate (batch_base_types.py:103-168) and the log-space primitives (util.py:17-47).
 *
 * instr: int32[n_instr][DFOL_INSTR_WORDS], question q executes instr[q_instr[q] .. q_instr[q+1]).
 * opts : int32 option columns (bit 30 = negated) referenced by terminal instructions.
 * lp_out: log-probabilities, written at instr.out (+k per option).
 * tape  : optional [n_instr][tape_stride] floats, the attention BEFORE each instruction (for backward).
 * Backward: d_lp (same indexing as lp_out) -> g_attr / g_rel: compact gradient slices w.r.t. the RAW table
 * entries each instruction read (offsets in instr words GA0/GA1/GR; slices of one option list are consecutive
 * with stride attr_stride[b] / rel_stride[b]); buffers must be zeroed by the caller.
 * mods (optional, may be NULL): float[rows][4] raw attention-transfer modulations (alpha/10, beta/10, c/10, d), one row
 * per predicate of every modulated sub-operator (BatchVariableSet.apply_modulations, batch_base_types.py:170-187;
 * rows indexed by the instruction words MOD / MOD2, +k per option); backward writes d loss / d mods into d_mods (same
 * shape, zeroed by the caller; every row has exactly one writer).
 */
#define DFOL_INSTR_WORDS 12
#define DFOL_I_OP 0
#define DFOL_I_FLAGS 1
#define DFOL_I_A0 2   /* concept / relation column, or first option index */
#define DFOL_I_A1 3   /* option count, or name column of a relate */
#define DFOL_I_A2 4   /* name column of choose_rel */
#define DFOL_I_OUT 5  /* index into lp_out */
#define DFOL_I_GA0 6  /* g_attr offset of the primary attribute operand (or of option 0) */
#define DFOL_I_GA1 7  /* g_attr offset of the name operand of relate / choose_rel */
#define DFOL_I_GR 8   /* g_rel offset of the relation operand (or of option 0) */
#define DFOL_I_MOD 9  /* attention-transfer modulation row of the primary sub-operator (of option 0), -1 = none */
#define DFOL_I_MOD2 10 /* ... of the name select of relate / choose_rel, or of the second branch (two_same, compare) */

#define DFOL_OP_SELECT 1
#define DFOL_OP_FILTER 2
#define DFOL_OP_RELATE 3
#define DFOL_OP_PUSH 4        /* end of the first branch: saved = cur */
#define DFOL_OP_EXIST 16      /* exist / end / tail of verify_rel */
#define DFOL_OP_AND 17
#define DFOL_OP_OR 18
#define DFOL_OP_VERIFY_ATTRS 19
#define DFOL_OP_CHOOSE_ATTR 20 /* choose_attr / query_attr */
#define DFOL_OP_CHOOSE_REL 21
#define DFOL_OP_ALL_SAME 22    /* flag NEGATE_RESULT -> all_different */
#define DFOL_OP_TWO_SAME 23    /* flag NEGATE_RESULT -> two_different */
#define DFOL_OP_COMPARE 24

#define DFOL_F_NEG 1            /* primary predicate is not(.) */
#define DFOL_F_ROUNDTRIP 2      /* some predicate of the op slot is negated: ll <- slog(exp(ll)) for the others */
#define DFOL_F_SUBJECT 4        /* relate: the new object is the subject */
#define DFOL_F_NAME_NEG 8
#define DFOL_F_NAME_ROUNDTRIP 16
#define DFOL_F_NORMALISE 32     /* softmax over the question's options (per object / per pair) */
#define DFOL_F_NEGATE_RESULT 64
#define DFOL_F_IS_LESS 128
#define DFOL_F_HARD 256         /* hard-mode quantifier (min instead of sum), eval only */
#define DFOL_OPT_NEG (1 << 30)

int dfol_program_fwd(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                     const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                     const float* rel_ll, const int64_t* rel_blk, const int32_t* rel_stride, const int32_t* img_n,
                     const float* mods, float* lp_out, float* tape, int tape_stride, void* stream);
int dfol_program_bwd(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                     const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                     const float* rel_ll, const int64_t* rel_blk, const int32_t* rel_stride, const int32_t* img_n,
                     const float* mods, const float* d_lp, const float* tape, int tape_stride, float* g_attr,
                     float* g_rel, float* d_mods, void* stream);

/* Tensor-core-mode builds of the two interpreter kernels (same contract, tape_stride = largest object count rounded
 * up to 4): MUFU exp/log approximations; the N x N tile of every relate hop is fetched with a bulk-async copy into a
 * ring of shared-memory buffers ahead of the hop that reads it, and the hop is evaluated in probability space:
 *   res[x] = prior[x] + slog(1 - prod_{y != x} (1 - p[x,y] e^{a[y]})),
 * a warp taking up to eight rows in one pass (a transposed shuffle butterfly reduces them together) and the lane that
 * ends up with a row writing its posterior.  rel_p (optional, NULL = absent): the PROBABILITY table p = e^{ll}, same
 * blocks / strides as rel_ll, zero on self pairs (written by dfol_rel_slots_fwd[_tc]); with it a pair costs one FFMA and
 * one FMUL, without it one MUFU.EX2 more.  rel_ll is still read by choose_rel and by programs beyond the ring's reach.
 * Results agree with the exact kernels to fp32 rounding of the approximations (bf16-mode tolerance of the answer
 * logits: 2e-2). */
int dfol_program_fwd_fast(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                          const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                          const float* rel_ll, const float* rel_p, const int64_t* rel_blk, const int32_t* rel_stride,
                          const int32_t* img_n, const float* mods, float* lp_out, float* tape, int tape_stride,
                          void* stream);
int dfol_program_bwd_fast(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                          const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                          const float* rel_ll, const float* rel_p, const int64_t* rel_blk, const int32_t* rel_stride,
                          const int32_t* img_n, const float* mods, const float* d_lp, const float* tape,
                          int tape_stride, float* g_attr, float* g_rel, float* d_mods, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training-mode dropout of the oracle networks (nn.Dropout in front of every Linear: nsvqa/nn/vision/regular_mlp.py:
 * 29-32, embedding_layer.py:73; sample_config.yaml: dropout 0.1).  The keep/drop decision of element (row, col) of
 * dropout site `site` is a pure function of (seed, site, row, col) (counter-based hash, 16 bits per element); kept elements
 * are scaled by 1/(1-p).  dfol_dropout_scale multiplies a row-major matrix in place (run it on a tensor of ones to
 * export a mask); dfol_pair_features_dropout writes the masked relation-network input rows
 * mask .* [obj_s | obj_o | geo(s,o)] of all pairs (batch_gqa_boxfeatures_pipeline.py:260-281), zero beyond 2*width+4.
 */
int dfol_dropout_scale(void* x, int64_t ld, int64_t rows, int cols, int is_bf16, uint64_t seed, int site, float p,
                       void* stream);
/* backward of the gather: d_obj[t, c] += sum_o dpm[(t,o), c] + sum_s dpm[(s,t), width + c] (dpm already masked; fp32 or
 * bf16) + addend_bf16[t, c] (optional: the attribute chain's share of d obj in tensor-core mode) */
int dfol_pair_features_bwd(const void* dpm, int64_t ld, int is_bf16, int width, float* d_obj, int64_t ldobj,
                           const void* addend_bf16, int64_t ld_add, const int32_t* pair_row, const int32_t* obj_row,
                           const int32_t* img_n, const int32_t* obj_img, int64_t objects, void* stream);
int dfol_pair_features_dropout(const float* obj, int64_t ldobj, int width, int pos_col, void* out, int64_t ldout,
                               int out_cols, int is_bf16, const int32_t* pair_row, const int32_t* obj_row,
                               const int32_t* img_n, const int32_t* pair_img, int64_t pairs, uint64_t seed, int site,
                               float p, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Attention-transfer calibrator, token side (BatchInterpreterBase.forward modulator loops, batch_base_interpreter.py:
 * 87-140; transform_attention of FilterBatch / RelateBatch, batch_base_ops.py:407-467, :598-684; _compute_attention_
 * modulations :275-286).  One LSTMCell evaluation over `rows` predicate rows, recurrent part only:
 *   pre = xproj[row] + b_hh + W_hh . (h_in[src] (+ h_add[src])),  src = owner ? owner[row] : row   (gates i, f, g, o)
 * xproj = features . W_ih^T + b_ih of ALL cells of the batch is one dfol_gemm_f32 by the caller.  mask (optional, 0/1 per
 * row): rows with 0 output the fallback state fb.  saved: [rows][7 S] (gates, c_in, tanh(c'), total h_in) for backward.
 * Backward writes d pre (rows x 4S), accumulates (+=) into the input / added / fallback state gradients; dW_ih, dW_hh and
 * the bias gradients are GEMM-shaped reductions over all cell rows done by the caller.  dfol_mod_out_*: the modulation
 * output layer sigmoid(W_out . [fh | bh] + b_out), its input kept in `cat`, and its backward (dzo = d pre-sigmoid).
 */
int dfol_lstm_cell_fwd(const float* xproj, int64_t ldx, const float* b_hh, const float* w_hh, int S, const float* h_in,
                       const float* c_in, const float* h_add, const float* c_add, const int64_t* owner,
                       const float* mask, const float* fb_h, const float* fb_c, float* h_out, float* c_out,
                       float* saved, int rows, void* stream);
int dfol_lstm_cell_bwd(const float* d_h_out, const float* d_c_out, const float* w_hh, int S, const float* saved,
                       const int64_t* owner, const float* mask, float* dpre, int64_t lddp, float* d_h_in,
                       float* d_c_in, float* d_h_add, float* d_c_add, float* d_fb_h, float* d_fb_c, int rows,
                       void* stream);
int dfol_mod_out_fwd(const float* fh, const float* bh, const int64_t* owner, const float* w_out, const float* b_out,
                     int S, int n_out, float* mods, float* cat, int rows, void* stream);
int dfol_mod_out_bwd(const float* d_mods, const float* mods, const int64_t* owner, const float* w_out, int S, int n_out,
                     float* dzo, float* d_fh, float* d_bh, int rows, void* stream);

/* The same passes as TWO persistent kernels (csrc/modulator_tape.cu): the caller compiles the sequence of steps of a
 * program batch into `records` (device array of n_rec records of dfol_mod_tape_record_size() bytes:
 *   { int32 kind (0 cell, 1 output layer, 2 squeeze, 3 gate), net, rows, base, live, pad;
 *     int64 in_h, in_c, add_h, add_c, fb_h, fb_c, out_h, out_c  -- offsets (floats) into the state pool, -1 = zero state;
 *     const int64* owner; const float* mask; const int64* part (row -> question, non-decreasing; NULL = identity) }),
 * both W_hh and the output layer stay in shared memory for the whole tape; the chain is sequential per question only, so
 * every block executes the tape for the rows of its own questions and dependent steps need a block barrier only.
 * pool / grad_pool: zero-filled by the caller; saved_* (R x 7S), cat (R x 2S), dpre_* (R x 4S), dzo (R x n_out) as in the
 * per-step kernels above (zero-filled: steps that do not run leave their rows untouched). */
int dfol_mod_tape_record_size(void);
int dfol_mod_tape_fwd(const void* records, int n_rec, int questions, float* pool, const float* xproj_f,
                      const float* xproj_b, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                      const float* b_hh_b, const float* w_out, const float* b_out, int S, int n_out, float* saved_f,
                      float* saved_b, float* mods, float* cat, void* stream);
int dfol_mod_tape_bwd(const void* records, int n_rec, int questions, float* grad_pool, const float* w_hh_f,
                      const float* w_hh_b, const float* w_out, int S, int n_out, const float* saved_f,
                      const float* saved_b, const float* mods, const float* d_mods, float* dpre_f, float* dpre_b,
                      float* dzo, void* stream);

/* dfol_pair_hidden_fwd_tc on the tensor cores: per 128-row tile the one-hot operand [e_s | e_o | geo | 1] is generated in
 * shared memory and multiplied by the image's own operand [U_b; V_b; Wg; bias] (gathered to bf16 into operand_workspace,
 * image_num * H * roundup(2 max_n + 5, 64) bf16), then ELU and the bf16 store.  max_n <= 125, H <= 256 (H % 16 == 0). */
int dfol_pair_hidden_fwd_mma(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos, const float* wg,
                             int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H, void* geo_out,
                             const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n, int image_num,
                             int max_n, void* operand_workspace, void* stream);
/* dfol_rel_slots_fwd on the tensor cores: a grouped GEMM, image b multiplies its activation rows by its own <= 16 slot
 * rows of W (gathered to bf16 into wb_workspace, 16 * image_num * K bf16) per pass; the activation (bf16, total_rows x K
 * valid columns, zero K padding) is read once whatever the number of slots.  Same tables and layout as dfol_rel_slots_fwd. */
int dfol_rel_slots_fwd_tc(const void* h_saved, int64_t ldh, int64_t total_rows, int E, int K, const float* W, int64_t ldw,
                          const float* bias, const int32_t* slot_wrow, const int32_t* img_slot, int max_slots,
                          const int64_t* slot_blk, const int32_t* stride, const int32_t* row0, const int32_t* img_rows,
                          const int32_t* img_n, int image_num, int max_rows, float diag_value, void* wb_workspace,
                          float* ll, float* p_out, void* stream);

/* Loss of VQATrainer._compute_loss (nsvqa/train/trainer.py:181-262) and its derivative w.r.t. lp.
 * kind 0 BINARY: BCE(exp(lp), target) summed; 1 QUERY: sum_q slog(sum_{k in q} e^{lp_k}) - sum_k target_k lp_k
 * (seg[q]..seg[q+1] are question q's predicates); 2 STATEMENT: -sum lp.  loss_out[0] += scale * loss,
 * d_lp = scale * dloss/dlp (scale = 1 / total question count, trainer.py:434-435). */
int dfol_loss_fwd_bwd(const float* lp, const float* target, const int32_t* seg, int n_seg, int n_lp, int kind,
                      float scale, float* loss_out, float* d_lp, void* stream);

/* Backward of the table (last) layer, consuming the compact gradient slices of dfol_program_bwd.
 * For slice j = (image b, table column c): dz[l] = g[goff[j] + l] * (1 - exp(LL[b][c][l])) (logsigmoid');
 * dH[(row0[b] + l), :] += dz[l] * W[wrow[j], :];  dW[wrow[j], :] += sum_l dz[l] * Hsaved[row0[b] + l, :];
 * db[wrow[j]] += sum_l dz[l].  Slices are grouped by image: image b owns slices [img_slice[b], img_slice[b+1]).
 * dH rows of an image are owned by one thread block (no atomics); dW/db use atomics.  All outputs zeroed by the
 * caller (dH must be zero for images without slices). */
int dfol_table_layer_bwd(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                         const int32_t* slice_wrow, const int32_t* img_slice, int image_num, const float* ll,
                         const int64_t* blk, const int32_t* stride, const int32_t* row0, const int32_t* img_rows,
                         const float* W, int64_t ldw, const float* h_saved, int64_t ldh, int E, float* dH,
                         int64_t lddh, float* dW, float* db, void* stream);

/* Fused variant for tables where an image touches at most 8 columns (the relation table): one pass over the rows
 * of every image, writes dZ = (sum_j dz_j W[wrow_j]) * act'(Hsaved) for ALL rows (zero where an image has no
 * slice: dZ needs no memset), accumulates dW / db with atomics.  max_rows = largest img_rows[b].
 * The caller guarantees img_slice[b+1] - img_slice[b] <= 8 for every image (otherwise use the general kernels).
 * bf16_io != 0: h_saved and dZ are bf16 (tensor-core mode); columns E <= e < out_cols of dZ are written as zero. */
int dfol_table_layer_bwd_fused(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                               const int32_t* slice_wrow, const int32_t* img_slice, int image_num, int max_rows,
                               const float* ll, const int64_t* blk, const int32_t* stride, const int32_t* row0,
                               const int32_t* img_rows, const float* W, int64_t ldw, const void* h_saved,
                               int64_t ldh, int E, int act, void* dZ, int64_t lddz, int out_cols, int bf16_io,
                               float* dW, float* db, void* stream);

/* Tensor-core-mode variant (bf16 h_saved / dZ, sigmoid below the table layer): any number of slices per image
 * (max_slices = largest img_slice[b+1]-img_slice[b]; consumed four per pass, later passes accumulate into dZ);
 * slice_col = column inside the image's table block, slice_wrow = row of W / dW / db; additionally
 * dbelow[e] += sum_rows dZ[row, e] (bias gradient of the layer below; NULL to skip).  Streams H and dZ once per pass
 * with 128-byte warp transactions. */
int dfol_table_layer_bwd_tc(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                            const int32_t* slice_wrow, const int32_t* img_slice, int image_num, int max_rows,
                            int max_slices, const float* ll, const int64_t* blk, const int32_t* stride,
                            const int32_t* row0, const int32_t* img_rows, const float* W, int64_t ldw,
                            const void* h_saved, int64_t ldh, int E, void* dZ, int64_t lddz, int out_cols, float* dW,
                            float* db, float* dbelow, float keep, void* stream);

/* The same backward on the tensor cores (csrc/table_layer_bwd_mma.cu; tensor-core mode, no dropout, 128 <= cols <= 320,
 * at most 32 slices per image): per 128-row tile of H (TMA) three tcgen05 contractions -- dZ = (DZ . Wslices) * h(1-h)
 * written in place and stored by TMA, dW^T += H^T . DZ accumulated per image in TMEM, dbelow += dZ^T . 1 accumulated per
 * CTA -- so H is read once and dZ written once.  tile_start[b] = prefix sums of ceil(img_rows[b] / 128)
 * (images + 1 entries, total_tiles = the last one); wb_workspace: images * 32 * cols bf16 (per-image slice rows of W). */
int dfol_table_layer_bwd_mma(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                             const int32_t* slice_wrow, const int32_t* img_slice, int image_num, int max_slices,
                             const float* ll, const int64_t* blk, const int32_t* stride, const int32_t* row0,
                             const int32_t* img_rows, const int32_t* tile_start, int total_tiles, int64_t total_rows,
                             const float* W, int64_t ldw, const void* h_saved, int64_t ldh, int E, void* dZ,
                             int64_t lddz, int cols, float* dW, float* db, float* dbelow, void* wb_workspace,
                             void* stream);

/* Pair-level forward chain in one kernel (csrc/pair_chain_fwd.cu; tensor-core mode, hidden width 256, no dropout):
 *   H1[(s,o), :] = elu(U[s] + V[o] + Wg . geo(s,o) + b1)   produced straight into the shared-memory A operand of
 *   H2 = sigmoid(H1 . W2^T + b2)                           the layer-2 tcgen05 GEMM (cluster of two CTAs, W2 resident)
 * (batch_gqa_boxfeatures_pipeline.py:257-279 + classifier_oracle.py:154).  uv = [U | V] fp32 (T, 2H), obj_pos = normalised
 * boxes, wg = geometry columns of the first Linear (ldw = its row stride), W2 bf16 [N, H]; pair_img[row] = image of a pair
 * row.  h1_out / geo_out (training: saved for the backward pass) may be NULL -- the forward pass itself never reads H1
 * from memory.  Results are bit-identical to dfol_pair_hidden_fwd_tc followed by dfol_pair_layer_fwd_cluster. */
int dfol_pair_chain_fwd(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos, const float* wg, int64_t ldw,
                        const float* bias1, const void* W2, int64_t ldw2, const float* bias2, void* H2, int64_t ldh2,
                        int store_cols, void* h1_out, int64_t ldh1, void* geo_out, const int32_t* pair_img,
                        const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n, int64_t M, int N, int H,
                        void* stream);

/* Demand-driven relation table (tensor-core mode; replaces computing all nR columns of
 * ClassifierOracle.compute_all_log_likelihood_2, classifier_oracle.py:154, when the batch's programs are known):
 * image b owns slots [img_slot[b], img_slot[b+1]); slot j evaluates row slot_wrow[j] of the embedding layer:
 * ll[slot_blk[b] + k*stride[b] + l] = logsigmoid(H[row0[b]+l, :] . W[wrow, :] + bias[wrow]) for the k-th slot of
 * the image, l < img_rows[b]; self pairs (l / img_n[b] == l %% img_n[b]) are written as diag_value. */
int dfol_rel_slots_fwd(const void* h_saved, int64_t ldh, int E, const float* W, int64_t ldw, const float* bias,
                       const int32_t* slot_wrow, const int32_t* img_slot, int max_slots, const int64_t* slot_blk,
                       const int32_t* stride, const int32_t* row0, const int32_t* img_rows, const int32_t* img_n,
                       int image_num, int max_rows, float diag_value, float* ll, float* p_out, void* stream);
/* p_out (optional, NULL = not written; also on dfol_rel_slots_fwd_tc): the probability table sigmoid(.) = e^{ll} in the
 * same layout, zero on self pairs -- the operand of the probability-space relate hop of dfol_program_{fwd,bwd}_fast. */

/* Dense variant for tables where images touch many columns (attribute options): scatters the slices into a zeroed
 * dense (rows x columns) matrix with logsigmoid' applied, dZ[row0[b] + l, col_j] += g_j[l] * (1 - exp(LL_j[l]));
 * the layer backward is then two ordinary GEMMs (dH = dZ.W, dW = dZ^T.H) and a column sum. */
int dfol_table_grad_dense(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                          const int32_t* slice_img, int slice_num, const float* ll, const int64_t* blk,
                          const int32_t* stride, const int32_t* row0, const int32_t* img_rows, float* dZ,
                          int64_t lddz, void* stream);

/* Answers on the device (eval / predict; reference: batch_gqa_ops.py:222-225, :404-407, :744-748, util.find_max_ind
 * util.py:64-66).  mode 0 binary (first = exp(lp) > 0.5), mode 1 option lists (segments seg[q]..seg[q+1]: the answer set
 * is every option whose probability equals the maximum exactly and exceeds `threshold`; first = first member or -1,
 * count = size, sel[k] = membership, may be NULL), mode 2 compare (argmax of the two entries).  best_lp = log-probability
 * of the first member. */
int dfol_answers(const float* lp, const int32_t* seg, int question_num, int mode, float threshold, int32_t* first,
                 int32_t* count, float* best_lp, uint8_t* sel, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimiser step on the flat parameter bucket: clip_grad_norm_ + Adam (trainer.py:438-441,
 * gqa_interpreter_experiments.py:261: torch.optim.Adam(lr, weight_decay) = L2 added to the gradient).
 * dfol_sumsq: out[0] += sum g^2.  dfol_adam_step: coef = min(1, clip / (sqrt(sumsq[0]) + 1e-6)) applied to g. */
int dfol_sumsq(const float* g, int64_t n, float* out, void* stream);
/* L1 regularisation of the loss (trainer.py:257-259, config key l1_lambda): g += coef * sign(p),
 * loss_out[0] += loss_coef * sum |p| (loss_out may be NULL).  Callers pass coef = lambda / (numel * batch). */
int dfol_l1_regularize(const float* p, float* g, int64_t n, float coef, float loss_coef, float* loss_out, void* stream);
int dfol_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq, float clip_norm,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFOL_B200_H */
