/* libdeepj_sm100.so -- C ABI of the B200-native DeepJ biaxial-LSTM hot path.
 *
 * Every entry point is what a binding for the reference's model.py /
 * generate.py hot path would call (reference = calclavia/music-generator;
 * file:line below are into that repository).  The reference has no native
 * code or FFI of its own: its "interface" for this path is the Keras layer
 * graph of model.py and the NumPy sampling loop of generate.py, so each
 * function names the graph segment it replaces.
 *
 * Conventions
 *   - plain C, no torch types; all pointers are DEVICE pointers on the current
 *     device unless the name says host; the caller owns every buffer.
 *   - every function returns 0 on success, <0 for an invalid argument /
 *     unsupported shape, >0 = cudaError_t.  dj_last_error() gives the text.
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as
 *     void*); no hidden synchronisation.
 *   - activations use one canonical row order: row = (b*T + t)*48 + n, so the
 *     reference's Permute layers (model.py:72,81,88) cost nothing.
 *   - there is NO CPU fallback.
 */
#ifndef DEEPJ_B200_H
#define DEEPJ_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DJ_NUM_NOTES 48      /* constants.py:54-56 */
#define DJ_NOTE_UNITS 3      /* constants.py:72 */
#define DJ_BEAT 16           /* constants.py:63 */
#define DJ_OCTAVE 12         /* constants.py:51 */
#define DJ_CONV_K 24         /* model.py:56 (2*OCTAVE) */
#define DJ_OCTAVE_UNITS 64   /* constants.py:70 */
#define DJ_STYLE_UNITS 64    /* constants.py:71 */
#define DJ_FEAT0 94          /* model.py:61-67: 1+12+1+64+16 */

/* dtype tags for GEMM operands produced by the glue kernels; DJ_F16 = IEEE half
 * (the 16-bit operand formats of tcgen05 kind::f16 are bf16 and half, per operand) */
#define DJ_F32 0
#define DJ_BF16 1
#define DJ_F16 2

/* One dropout site (model.py:58,80,85,116,123,136-138).  mode 0 = inactive
 * (predict), 1 = one hash word per 4 elements (rate*256 integral), 2 = one
 * word per element.  Built by dj_make_dropout.  key_ptr (nullable, DEVICE
 * pointer): when set, the kernels read the site key from it instead of `key` --
 * the per-step seed then lives in device memory and a captured CUDA graph of the
 * training step can be replayed with a new seed every step. */
typedef struct dj_dropout {
  uint32_t key;
  uint32_t thr;
  float scale;
  int32_t mode;
  const uint32_t* key_ptr;
} dj_dropout;
/* host helper: the site key dj_make_dropout derives from (seed, site) */
uint32_t dj_dropout_site_key(uint64_t seed, int site);

int dj_version(void);
const char* dj_last_error(void);
/* Deterministic gradients.  The gradient kernels that sum over CTAs (dj_wgrad_gemm_*, dj_lstm_scan_tc_bwd's db,
 * dj_conv_bwd, dj_colsum, split-K dj_gemm_simt) add with fp32 atomics by default, so two runs agree only to rounding.
 * Register a device workspace of `floats` floats for a stream (per calling thread; floats = 0 unregisters) and the
 * launches on that stream write one partial per CTA there and add them in index order instead: bit-identical
 * results run to run (also dj_lstm_scan_bwd's db).  A registered workspace too small for a launch makes that launch
 * fail (<0), it never falls back to the atomics.  8 M floats cover the default and the scaled model at any batch. */
int dj_set_reduce_workspace(void* stream, float* ws, int64_t floats);
/* host helper: fills *out for Dropout(rate) at `site` (1..12 = D1..D12 of SURVEY 8a). */
int dj_make_dropout(uint64_t seed, int site, float rate, dj_dropout* out);
/* debug: dump the 0/1 keep mask of a site with `rows` rows of F elements
 * (element index = row*roundup4(F)+f) so a CPU checker can replay it. */
int dj_dropout_mask_materialize(dj_dropout d, int64_t rows, int F, float* out, void* stream);

/* ---- front end --------------------------------------------------------------
 * dj_style_fwd replaces Dense(64,name='style') (model.py:141-142) and the four
 * per-layer Dense(F)+tanh style projections (model.py:77-79,113-115).
 *   style_in[b*style_bstride + t*style_tstride + s], s<num_styles
 *   emb [B*T,64];  sp[l] [B*T,F[l]] = tanh(emb.Wsd[l]+bsd[l]) */
int dj_style_fwd(const float* style_in, int64_t style_bstride, int64_t style_tstride, int num_styles,
                 int B, int T, const float* Ws, const float* bs, int n_proj,
                 const float* const* Wsd, const float* const* bsd, const int* F,
                 float* emb, float* const* sp, void* stream);
/* dj_frontend_fwd replaces input Dropout (model.py:136-137), the octave
 * Conv1D+tanh+Dropout (model.py:56-58), pitch_pos/pitch_class/pitch_bins
 * (model.py:22-49, including the reshape scramble over the LOCAL batch), the
 * beat RepeatVector, Concatenate, Permute (model.py:61-72) and the layer-0
 * style Add (model.py:78-82).  Writes the time-axis layer-0 GEMM A operand
 * A0[M, ldA] (cols >= 94 zero).  A0_lo (nullable, DJ_BF16 only) receives the
 * bf16 residual A - bf16(A): the second operand of the split gate GEMM. */
int dj_frontend_fwd(const float* notes_in, int64_t notes_bstride, const float* beat_in,
                    int64_t beat_bstride, int B, int T, const float* Wc, const float* bc,
                    const float* sp0, dj_dropout d_notes, dj_dropout d_beat, dj_dropout d_conv,
                    dj_dropout d_sp, void* A0, void* A0_lo, int ldA, int a_dtype, void* stream);
/* dj_layer_input replaces, for every LSTM layer but the first, the Dropout of
 * the previous layer output (model.py:85,123), the style Add (model.py:82,117)
 * and -- for note-axis layer 0 -- shift_chosen + Concatenate (model.py:101-106,
 * chosen post input-Dropout model.py:138).
 *   h_prev row for (b,t,n) = h_row0 + b*h_b_rows + t*48 + n, Uprev columns
 *   chosen_in (nullable) [B,T,48,3] with batch stride chosen_bstride
 *   A[M, ldA]: cols [0,Uprev) = drop(h) + drop(sp); cols [Uprev,F) = shifted
 *   chosen + drop(sp); cols >= F zero.  A_lo as in dj_frontend_fwd. */
int dj_layer_input(const float* h_prev, int Uprev, int64_t h_row0, int64_t h_b_rows, dj_dropout d_h,
                   const float* sp, int F, dj_dropout d_sp, const float* chosen_in,
                   int64_t chosen_bstride, dj_dropout d_chosen, int B, int T, void* A, void* A_lo, int ldA,
                   int a_dtype, void* stream);

/* ---- gate projections (the dense part of keras LSTM: x.W + b, model.py:84,120)
 * Generic CUDA-core GEMM with fp32 accumulation; operands DJ_F32 or DJ_BF16
 * (generation path at fp32, small style/conv gradients, cross-check of the
 * tensor-core kernels):
 *   C[m,n] (+)= sum_k A(m,k) B(k,n) (+ bias[n])
 *   A(m,k) = A[m*a_sm + k*a_sk], B(k,n) = B[k*b_sk + n*b_sn], C[m*ldc + n]
 *   a_shift/a_period: if a_period>0, A(m,k) reads row k-a_shift and is zero
 *   where (k % a_period) < a_shift  (h_{t-1} operand of the U weight gradient). */
int dj_gemm_simt(const void* A, int a_dtype, int64_t a_sm, int64_t a_sk, const void* B, int b_dtype,
                 int64_t b_sk, int64_t b_sn, float* C, int64_t ldc, const float* bias, int M, int N,
                 int K, int accumulate, int64_t a_shift, int64_t a_period, void* stream);
/* tcgen05 + TMA bf16 GEMM, fp32 accumulate in TMEM:
 *   C[M,N] = A[M,K] . Bt[N,K]^T + bias[N]      (both operands K-major, bf16)
 *   lda/ldb in elements (multiples of 8), K padded with zeros up to lda. */
int dj_gate_gemm_bf16(const void* A, int64_t lda, const void* Bt, int64_t ldb, float* C, int64_t ldc,
                      const float* bias, int M, int N, int K, void* stream);
/* The same kernel with a choice of 16-bit format (DJ_BF16 / DJ_F16, the SAME for both
 * operands: tcgen05 kind::f16 raises an illegal-instruction fault on B200 when one
 * operand is half and the other bf16, so a_fmt != b_fmt is rejected) and, when
 * A_lo and Bt_lo are given, the fp32-grade SPLIT product in three tcgen05 passes
 * over the same TMEM accumulator:
 *   C = A.Bt^T + A_lo.Bt^T + A.Bt_lo^T + bias,   X_lo = 16-bit(X - 16-bit(X))
 * (the lo.lo term, 2^-16 relative, is dropped).  The forward x.W+b of training
 * runs this way: north_star's 1e-3 tolerance on the outputs is not reachable
 * with single bf16 operands (tools/precision_study.py). */
int dj_gate_gemm_16(const void* A, const void* A_lo, int a_fmt, int64_t lda, const void* Bt,
                    const void* Bt_lo, int b_fmt, int64_t ldb, float* C, int64_t ldc, const float* bias,
                    int M, int N, int K, void* stream);
/* ... and with the accumulator multiplied by out_scale before the bias is added: the generation path splits
 * operands into IEEE half hi+lo (22 mantissa bits) after scaling the weights by a power of two, which keeps the
 * residuals out of half's subnormal range; out_scale undoes it exactly. */
int dj_gate_gemm_16s(const void* A, const void* A_lo, int a_fmt, int64_t lda, const void* Bt,
                     const void* Bt_lo, int b_fmt, int64_t ldb, float* C, int64_t ldc, const float* bias,
                     float out_scale, int M, int N, int K, void* stream);
/* tcgen05 weight-gradient GEMM (contraction over the M rows, split across CTAs,
 * fp32 global reductions):  C[Ka,Nb] += A[M,Ka]^T . B[M,Nb]   (bf16, MN-major).
 * The h_{step-1} operand of dU comes pre-shifted from dj_lstm_scan_fwd. */
int dj_wgrad_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                       int Ka, int Nb, int64_t M, void* stream);
/* the same with a choice of format (a_fmt == b_fmt, see dj_gate_gemm_16) */
int dj_wgrad_gemm_16(const void* A, int a_fmt, int64_t lda, const void* B, int b_fmt, int64_t ldb, float* C,
                     int64_t ldc, int Ka, int Nb, int64_t M, void* stream);
/* In-place IEEE half -> bf16 of n elements.  The forward scan keeps h_{step-1} in half
 * (4 more mantissa bits in the recurrence); its weight gradient dU = H_{step-1}^T.dZ
 * multiplies it with the bf16 dZ, so after the forward pass the buffer is converted. */
int dj_half_to_bf16_inplace(void* buf, int64_t n, void* stream);
/* fp32 -> bf16 operand copies: out[r, c] = in[r, c] (c<cols) else 0, out ld = ldo;
 * transpose!=0 writes out[c, r] (ldo >= rows). */
int dj_cast_bf16(const float* in, int rows, int cols, void* out, int ldo, int transpose, void* stream);
/* the same for up to 16 tensors in one launch (all operand copies of a training step):
 * fmt[i] = DJ_BF16 / DJ_F16; out_lo (nullable array, nullable entries) receives the
 * 16-bit residual in - 16-bit(in) in the same layout. */
int dj_cast16_multi(int n, const float* const* in, const int* rows, const int* cols, void* const* out,
                    void* const* out_lo, const int* ldo, const int* transpose, const int* fmt,
                    const float* scale /* nullable: per-entry power-of-two pre-scale */, void* stream);

/* ---- recurrence (the sequential part of keras LSTM, model.py:84,120) ---------
 * Persistent thread-block-cluster kernel: the recurrent weights U [units,4*units]
 * stay in shared memory, split by hidden unit over the CTAs of a cluster; h is
 * exchanged through distributed shared memory; cell state lives in registers.
 *   Z [M,4*units] fp32: in = x.W+b, out = activated gates i,f,g,o (saved for bwd);
 *     gate columns are GATE-INTERLEAVED: col = 4*unit + gate (the host converts
 *     Keras' [i|f|c|o] block order at the API boundary); same for U, dZ, db
 *   row(seq, step) = (seq/seq_inner)*seq_outer_stride + (seq%seq_inner)*seq_inner_stride
 *                    + step*step_stride
 *   h_out [M,units] fp32; c_out nullable (training only); h_prev_bf16 nullable:
 *   bf16 [M,units] receiving h_{step-1} at each row (zeros at step 0) = the A
 *   operand of the recurrent weight gradient dU = H_{step-1}^T.dZ.
 *   hard != 0: hard_sigmoid gates (Keras < 2.3), else sigmoid. */
int dj_lstm_scan_fwd(float* Z, float* h_out, float* c_out, void* h_prev_bf16, const float* Uw, int S,
                     int steps, int units, int seq_inner, int64_t seq_outer_stride,
                     int64_t seq_inner_stride, int64_t step_stride, int hard, void* stream);
/* Tensor-core variant of the forward recurrence (training path, 16-bit operands for
 * h.U, fp32 accumulate/state): same row maps as dj_lstm_scan_fwd, but Z is read only and the
 * activated gates i,f,g,o are saved to gates16 [M, units] x 4 IEEE halves (8 bytes per cell,
 * the input of dj_lstm_scan_tc_bwd; a hard-sigmoid gate strictly inside (0,1) is stored
 * strictly inside, so the derivative's indicator survives the rounding); U is passed as
 * Ut_16 [4*units, units] (transposed, gate-interleaved rows) in format `fmt`
 * (DJ_BF16 / DJ_F16), h_prev_16 (same format) is REQUIRED (it doubles as the
 * inter-CTA exchange buffer), and Ut_lo (nullable) is the 16-bit residual
 * U^T - fmt(U^T): with it every step runs a second MMA pass h.U_lo, so the recurrent
 * weights enter with ~22 mantissa bits and only h is rounded (to half: 2^-12).
 * Only the two maps of the model are supported: units 256/512 with the time-axis
 * map, units 128/256 with the note-axis map. */
int dj_lstm_scan_tc_fwd(const float* Z, void* gates16, float* h_out, float* c_out, void* h_prev_16,
                        const void* Ut_16, const void* Ut_lo, int fmt, int S, int steps, int units, int seq_inner,
                        int64_t seq_outer_stride, int64_t seq_inner_stride, int64_t step_stride, int hard,
                        void* stream);
/* Inference variant for the generation window (generate.py:106-109: both time-axis layers over
 * the 128-step window from zero state, every generated timestep).  The sampled events must equal
 * the fp32 model's, so h.U is fp32-grade: U as half hi + lo (pre-scaled by 1/acc_scale, a power of
 * two, so the residual stays normal), h_{t-1} as half hi + lo exchanged through h_hi / h_lo
 * [M, units], three MMA passes per step (U_hi.h_hi + U_lo.h_hi + U_hi.h_lo, fp32 accumulators in
 * tensor memory).  Z is read only; nothing is saved for a backward pass.  Time-axis map, 256 units. */
int dj_lstm_scan_tc_infer(const float* Z, float* h_out, void* h_hi, void* h_lo, const void* Ut_hi,
                          const void* Ut_lo, float acc_scale, int S, int steps, int units, int seq_inner,
                          int64_t seq_outer_stride, int64_t seq_inner_stride, int64_t step_stride, int hard,
                          void* stream);
/* Generation with one or two sequences: BOTH time-axis layers (model.py:75-85 at inference, generate.py:106-109) in
 * one launch, layer 1 running one step behind layer 0 (129 sequential step latencies per window instead of 256).
 * Layer 1's input projection is part of its recurrent product: with no dropout its input is h0_t + sp (sp = the style
 * projection, constant over the window), so z1_t = h0_t.W1 + c1 + h1_{t-1}.U1 with c1[b, 4U] = sp[b].W1 + b1
 * computed by the caller (fp32, gate-interleaved columns like Z).  Z0 [S*steps, 4U] = layer 0's x.W + b.
 * h0_hi/h0_lo/h1_hi/h1_lo: the layers' exchange buffers, IEEE half, [S*steps + 48, U] rows (one spare timestep).
 * Weights as for dj_lstm_scan_tc_infer (half hi + lo, pre-scaled by 1/acc_scale); Wt1 = layer 1's W^T [4U, 256].
 * One sequence (S = 48) runs as THREE sets of clusters -- layer 0, the input projection of layer 1 (which writes
 * z1in_t = h0_t.W1 + c1 into Z1 [S*steps, 4U], scratch) and layer 1 --, two sequences as two (projection fused into
 * layer 1's MMA; Z1 unused).  flags: 2*S/16 scratch words.  h1_out [S*steps, U]: only the rows of the LAST step are
 * written (time_out[:, -1]).
 * Time-axis map (seq = (b, n), rows (b*steps + t)*48 + n), 256 units, S <= 96. */
int dj_lstm_scan_tc_gen2(const float* Z0, float* Z1, const float* c1, float* h1_out, void* h0_hi, void* h0_lo, void* h1_hi,
                         void* h1_lo, const void* Ut0_hi, const void* Ut0_lo, const void* Ut1_hi, const void* Ut1_lo,
                         const void* Wt1_hi, const void* Wt1_lo, float acc_scale, uint32_t* flags, int S, int steps,
                         int hard, void* stream);
/* Tensor-core variant of the reverse scan (dz.U^T on tcgen05): U is passed as
 * Un_bf16 [units, 4*units] (bf16, natural, gate-interleaved columns); dZ is bf16
 * (gradients need its exponent range) and doubles as the inter-CTA exchange buffer;
 * db accumulates with fp32 atomics. */
int dj_lstm_scan_tc_bwd(const void* gates16, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                        const void* Un_bf16, void* dZ_bf16, float* db, int S, int steps, int units,
                        int seq_inner, int64_t seq_outer_stride, int64_t seq_inner_stride,
                        int64_t step_stride, int hard, void* stream);
/* reverse scan: consumes gates/c and dY (gradient w.r.t. the DROPPED-OUT layer
 * output, row stride ldY; the kernel applies the mask d_y itself), produces
 * dZ [M,4*units] (dz_dtype) and accumulates db[4*units] (fp32 atomics). */
int dj_lstm_scan_bwd(const float* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                     const float* Uw, void* dZ, int dz_dtype, float* db, int S, int steps, int units,
                     int seq_inner, int64_t seq_outer_stride, int64_t seq_inner_stride,
                     int64_t step_stride, int hard, void* stream);

/* ---- heads + loss (model.py:94-95,125 and primary_loss model.py:14-20) -------
 * probs[M,3] = [sigmoid(x.Wn+bn), x.Wv+bv], x = drop(h).  With y_true != NULL it
 * also writes per-block loss / head-gradient partials and dX[M,units] (gradient
 * w.r.t. x).  partials must hold dj_head_partials_size(units) floats. */
int64_t dj_head_partials_size(int units);
int dj_head_loss(const float* h, int units, dj_dropout d_h, const float* Wn, const float* bn,
                 const float* Wv, const float* bv, const float* y_true, float* probs, float* dX,
                 float* partials, int64_t M, void* stream);
/* sums the partials: loss_out[0] = loss; dWn[units,2], dbn[2], dWv[units], dbv[1] (overwrite). */
int dj_head_finalize(const float* partials, int units, float* loss_out, float* dWn, float* dbn,
                     float* dWv, float* dbv, void* stream);

/* ---- backward glue ----------------------------------------------------------
 * ds[bt,f] = (1-sp^2) * sum_n drop_mask(d_sp)(row,f) * dA[row,f]   (model.py:78-82 backward) */
int dj_style_bwd_reduce(const float* dA, int64_t ldA, int F, const float* sp, dj_dropout d_sp,
                        int BT, float* ds, void* stream);
/* out[c] (+)= sum_r X[r*ldx + c] */
int dj_colsum(const float* X, int64_t ldx, int64_t R, int C, float* out, int accumulate, void* stream);
/* conv weight/bias gradient (model.py:56-58 backward); recomputes tanh(conv);
 * dA0 = gradient of the layer-0 A operand [M, ldA] fp32; accumulates (atomics). */
int dj_conv_bwd(const float* notes_in, int64_t notes_bstride, int B, int T, const float* Wc,
                const float* bc, dj_dropout d_notes, dj_dropout d_conv, const float* dA0, int64_t ldA,
                float* dWc, float* dbc, void* stream);

/* ---- optimizer: keras.optimizers.Nadam (model.py:152) on flat buffers --------
 * g is scaled by gscale (1/world for the NCCL-summed gradient) first. */
int dj_nadam_step(float* p, const float* g, float* m, float* v, int64_t n, float gscale, float lr,
                  float beta1, float beta2, float eps, float mu_t, float mu_t1, float m_sched_new,
                  float m_sched_next, float bias2, void* stream);

/* The same update with the ten per-step scalars read from DEVICE memory (graph replay):
 * sc = {gscale, lr, beta1, beta2, eps, mu_t, mu_t1, 1/(1-m_sched_new), 1/(1-m_sched_next), 1/bias2}. */
int dj_nadam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* sc, void* stream);

/* ---- data-parallel step: gradient exchange fused with Nadam over peer memory ----
 * The reference trains on one device (train.py:29 model.fit); SURVEY.md 8e adds
 * data parallelism with one process per GPU.  Instead of an all-reduce followed by
 * dj_nadam_step, every rank launches ONE kernel that waits for all local gradients,
 * reduces its 1/world slice of the flat gradient with loads from every rank's
 * buffer (fixed rank order), applies Nadam to that slice (the moments of a slice
 * live on its owner only) and stores the new weights into every rank's parameter
 * buffer, then waits until all ranks are done.
 *   dj_peer_alloc   cudaMalloc'd, zeroed buffer + its 64-byte CUDA IPC handle
 *   dj_peer_open    map another rank's buffer (same node, peer access over NVLink)
 *   dj_peer_flag_words  size, in uint32 words, of the per-rank flag block (zeroed)
 *   peer_params / peer_grads / peer_flags: host arrays [world] of the ranks' buffers
 *   in THIS process' address space (entry `rank` is the local one); n floats, n%4==0;
 *   epoch = 1, 2, 3, ... the same on every rank.  Waits are bounded (20 s): on a
 *   timeout word [dj_peer_flag_words()-1] of the local flag block is set to 1. */
int64_t dj_peer_flag_words(void);
int dj_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int dj_peer_open(const unsigned char* handle64, void** ptr);
int dj_peer_close(void* ptr);
int dj_peer_free(void* ptr);
int dj_nadam_allreduce_peer(float* const* peer_params, const float* const* peer_grads,
                            uint32_t* const* peer_flags, int rank, int world, float* m, float* v,
                            int64_t n, uint32_t epoch, float gscale, float lr, float beta1, float beta2,
                            float eps, float mu_t, float mu_t1, float m_sched_new, float m_sched_next,
                            float bias2, void* stream);

/* ---- generation (generate.py:47-79,98-121) -----------------------------------
 * Persistent note-by-note sampler: for each of `G` sequences walks the 48 notes
 * carrying the note-axis LSTM state, applies the temperature transform, draws
 * play/replay against the uniform stream and writes the event row.
 *   zpre [G*48, 4*units]: note-layer-0 projection of [time_out + style, style]
 *        (everything but the chosen_{n-1} term), from the gate GEMM
 *   W0c [3,4*units]: rows Ut..Ut+2 of note0.lstm.W;  U0, W1, U1, b1 as in Keras
 *   sp1 [G,units]: tanh style projection of note layer 1
 *   stream_mode 0: reference stream -- one CTA walks all sequences, uniforms
 *        consumed n-major/sequence-minor with the conditional second draw,
 *        *ucursor (device int64) advanced;  1: indexed stream U[g][n][2].
 *   state per sequence: temperature (double), silent_time (int32)
 *   events [G,48,3] written (also appended by the caller to the window),
 *   probs_out nullable [G,48,3], margin_out nullable [G] (min |u-p|). */
int dj_gen_sample(const float* zpre, const float* W0c, const float* U0, const float* W1,
                  const float* U1, const float* b1, const float* sp1, const float* Wn,
                  const float* bn, const float* Wv, const float* bv, int units, int G,
                  const double* uniforms, int64_t* ucursor, int stream_mode, double* temperature,
                  int32_t* silent_time, double default_temp, int hard, float* events,
                  float* probs_out, double* margin_out, void* stream);

/* graph-replay form: epoch and the ten scalars of dj_nadam_step_dev come from device memory */
int dj_nadam_allreduce_peer_dev(float* const* peer_params, const float* const* peer_grads,
                                uint32_t* const* peer_flags, int rank, int world, float* m, float* v,
                                int64_t n, const uint32_t* epoch_dev, const float* sc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPJ_B200_H */
