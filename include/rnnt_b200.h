/* rnnt_b200.h -- C ABI of the B200-native RNN-T transducer head.
 *
 * Drop-in boundary for the one hot path this repository accelerates: the RNN-T joint network
 * (f_t + g_u -> tanh -> Linear(H,V) -> log_softmax) fused with the RNNTLoss alpha/beta lattice and
 * its gradient, plus the greedy-decode joint step.  The reference (MyrtleSoftware/myrtlespeech) is
 * pure Python on stock PyTorch and its snapshot contains no RNN-T code, so each entry
 * point cites the in-tree CTC analog whose role it takes:
 *
 *   rnnt_fused_forward / rnnt_fused_backward
 *       loss module forward + autograd:  src/myrtlespeech/loss/ctc_loss.py:51-101 (LogSoftmax at :95,
 *       library loss at :96-101), called from src/myrtlespeech/run/train.py:67 and :73; the Linear it
 *       absorbs is src/myrtlespeech/model/fully_connected.py:118-126,164.
 *   rnnt_lattice_forward
 *       the loss alone on materialised log-probabilities (same call sites), for callers that already
 *       hold a (B,T,U+1,V) tensor.
 *   rnnt_greedy_decode_lstm
 *       the decoder's whole frame loop, src/myrtlespeech/post_process/ctc_greedy_decoder.py:77-92, with the recurrent
 *       step src/myrtlespeech/model/rnn.py:133-205 of the prediction network inside it.
 *   rnnt_greedy_joint_argmax
 *       the per-step argmax of src/myrtlespeech/post_process/ctc_greedy_decoder.py:74, called from
 *       src/myrtlespeech/run/run.py:94.
 *
 * Conventions: plain pointers and sizes only; every device buffer (inputs, outputs, workspace) is
 * allocated by the caller; calls are stream-ordered on `stream` (a cudaStream_t passed as void*), never
 * synchronise the host, never allocate and keep no global mutable state (the rnnt_debug_* knobs at the end of
 * this header are process-wide test / bring-up switches and the only exception).  Return value 0 = success;
 * otherwise one of RNNT_ERR_* and rnnt_last_error() describes it.  Length arrays are HOST int32 (they
 * originate on the host: src/myrtlespeech/data/batch.py:103-105); label ids are DEVICE int32
 * (src/myrtlespeech/builders/task_config.py:103-108).
 *
 * Layouts (row-major, contiguous):
 *   f   bf16 [B][Tmax][H]        encoder output        g   bf16 [B][Umax+1][H]   prediction output
 *   W   bf16 [V][H]              joint projection      bias f32 [V] (may be NULL)
 *   y   i32  [B][Umax]           label ids (never `blank`)
 *   loss f32 [B]                 -ln P(y_b | x_b)
 *   df  f32 [B][Tmax][H]   dg f32 [B][Umax+1][H]   dW f32 [V][H]   db f32 [V]
 * Constraints: H % 8 == 0, 1 <= V <= 8192, Umax + 1 <= 4096, 1 <= f_lens[b] <= Tmax, 0 <= y_lens[b] <= Umax.
 */
#ifndef RNNT_B200_H_
#define RNNT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNNT_OK 0
#define RNNT_ERR_INVALID_ARGUMENT 1
#define RNNT_ERR_WORKSPACE_TOO_SMALL 2
#define RNNT_ERR_CUDA 3
#define RNNT_ERR_UNSUPPORTED 4

/* ABI version of this header (bumped on any signature change). */
int rnnt_abi_version(void);

/* Human-readable description of the last non-zero return on this thread. */
const char* rnnt_last_error(void);

/* Bytes of device workspace the fused calls need for these maxima.  Pure host arithmetic. */
size_t rnnt_fused_workspace_bytes(int B, int Tmax, int Umax, int V, int H);

/* Only the first rnnt_fused_state_bytes(...) bytes of the workspace carry information from rnnt_fused_forward to
 * rnnt_fused_backward (lengths, row logsumexp, lp_blank / lp_label, arc occupancies: ~40 MB at B=32 T=500 U=100);
 * the rest (~0.4 GB) is scratch.  A caller that keeps several forward passes alive may save just that prefix per
 * pass and copy it to the front of any workspace of the right size before the matching backward call. */
size_t rnnt_fused_state_bytes(int B, int Tmax, int Umax, int V, int H);

/* Joint + log-softmax + alpha/beta.  Writes loss[B]; leaves lse / lp_blank / lp_label / arc
 * occupancies in `workspace` for rnnt_fused_backward.  The B*T*(U+1)*V logits are never written. */
int rnnt_fused_forward(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                       const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax,
                       int V, int H, int blank, float* loss, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Gradient of sum_b grad_loss[b] * loss[b].  `workspace` must be the one the matching forward filled.
 * Overwrites df, dg, dW, db. */
int rnnt_fused_backward(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                        const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax,
                        int V, int H, int blank, const float* grad_loss, float* df, float* dg, float* dW,
                        float* db, void* workspace, size_t workspace_bytes, void* stream);

/* The same pair with the joint's activations KEPT between the calls.  rnnt_fused_forward above never writes the
 * B*T*(U+1)*V logits and rnnt_fused_backward recomputes h = tanh(f + g) and the logits (2 of the 8 N*H*V flops of a step
 * and N*H tanh); with 180 GB of HBM per GPU the other trade is usually the better one: the forward call also writes the
 * logits (fp16, base-2, bias included) and h (bf16) of every lattice row into `kept`, a device buffer of
 * rnnt_fused_kept_bytes(...) bytes owned by the caller (6.8 GB at B=32 T=500 U=100 V=H=1024), and the backward call
 * streams them back instead of running the recompute GEMM.  The buffer must stay untouched between the two calls; it is
 * the only state outside `workspace` (several live forward passes need one buffer each).  rnnt_fused_kept_bytes returns 0
 * where nothing can be kept (more than 4096 vocabulary columns); a NULL / too small buffer, or a shape the one-launch
 * backward kernel does not cover, silently selects the recompute schedule.  A non-NULL `kept` passed to
 * rnnt_fused_backward_kept MUST be the buffer the matching rnnt_fused_forward_keep call filled (same shapes, same lengths):
 * the library cannot tell a filled buffer from an unfilled one.  Gradients differ from the recompute schedule
 * only through the fp16 rounding of the kept logits (relative 2^-11 on a logit, below the bf16 rounding of dz that both
 * schedules share); the blank and label columns use the exact fp32 log-probabilities in both. */
size_t rnnt_fused_kept_bytes(int B, int Tmax, int Umax, int V, int H);
int rnnt_fused_forward_keep(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                            const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax,
                            int V, int H, int blank, float* loss, void* workspace, size_t workspace_bytes,
                            void* kept, size_t kept_bytes, void* stream);
int rnnt_fused_backward_kept(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                             const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax,
                             int V, int H, int blank, const float* grad_loss, float* df, float* dg, float* dW,
                             float* db, void* workspace, size_t workspace_bytes, const void* kept,
                             size_t kept_bytes, void* stream);

/* Loss on materialised log-probabilities: lp_blank, lp_label f32 [B][Tmax][Umax+1] (natural layout,
 * lp_label[b][t][u] = log p(y[b][u] | t,u), column Umax unused).  Writes loss[B] and the arc
 * occupancies c_blank, c_label f32 [B][Tmax][Umax+1] with
 *   d loss_b / d logits[b,t,u,k] = softmax_k (c_blank + c_label) - [k==blank] c_blank - [k==y_u] c_label. */
size_t rnnt_lattice_workspace_bytes(int B, int Tmax, int Umax);
int rnnt_lattice_forward(const float* lp_blank, const float* lp_label, const int32_t* f_lens_host,
                         const int32_t* y_lens_host, int B, int Tmax, int Umax, float* loss, float* c_blank,
                         float* c_label, void* workspace, size_t workspace_bytes, void* stream);

/* One greedy-decode joint step for B utterances: out_k[b] = argmax_v (W . tanh(f[b][t_idx[b]] + g[b]) + bias)
 * (lowest index wins ties), or -1 where t_idx[b] < 0.  g is bf16 [B][H]; t_idx, out_k are DEVICE int32 [B]. */
int rnnt_greedy_joint_argmax(const void* f, const void* g, const void* W, const float* bias,
                             const int32_t* t_idx, int32_t* out_k, int B, int Tmax, int V, int H, void* stream);

/* One whole greedy decode step for B utterances, bookkeeping included (the loop body of
 * src/myrtlespeech/post_process/ctc_greedy_decoder.py:77-92 without its per-frame host sync).  For every utterance with
 * t_cur[b] < lens[b]: k = argmax_v (W . tanh(f[b][t_cur[b]] + bf16(g[b])) + bias); if k != blank it is appended to
 * sym[b][n_sym[b]++] and emitted[b] is incremented; the frame advances (t_cur[b]++, emitted[b] = 0) on blank or once
 * max_symbols were emitted at this frame.  is_sym[b] / label[b] tell the caller where to step the prediction network
 * (label = 0 where nothing was emitted); active[b] = t_cur[b] < lens[b] after the update.  g is f32 [B][H]; all state
 * arrays are DEVICE int32 and updated in place. */
int rnnt_greedy_step(const void* f, const float* g, const void* W, const float* bias, const int32_t* lens,
                     int32_t* t_cur, int32_t* emitted, int32_t* n_sym, int32_t* sym, int sym_cap, int32_t* is_sym,
                     int32_t* label, int32_t* active, int B, int Tmax, int V, int H, int blank, int max_symbols,
                     void* stream);

/* The whole greedy decode of a batch in ONE launch: prediction-network LSTM cell (single layer) + output projection +
 * joint argmax + the emit / advance bookkeeping of rnnt_greedy_step, looped on the device until every utterance has
 * consumed its frames (replaces the per-frame host loop of src/myrtlespeech/post_process/ctc_greedy_decoder.py:77-92
 * and the recurrent step of src/myrtlespeech/model/rnn.py:133-205 inside it).
 *   f          bf16 [B][Tmax][H]   encoder output           lens      DEVICE i32 [B]
 *   W, bias    joint projection as above
 *   gate_table f32 [V+1][4*Hp]     W_ih . emb[v] + b_ih + b_hh for every label v (torch gate order i,f,g,o);
 *                                  row V is the start-of-sequence input
 *   W_hh       bf16 [4*Hp][Hp]     recurrent weights (torch layout)
 *   W_proj     bf16 [H][Hp], bias_proj f32 [H] (may be NULL): prediction output g = W_proj . h + bias_proj
 *   sym        DEVICE i32 [B][sym_cap] emitted ids, n_sym DEVICE i32 [B] their count (sym_cap >= Tmax*max_symbols holds all)
 * Arithmetic: h and W_* enter the tensor cores as bf16 with fp32 accumulation; the cell state c and the gate
 * non-linearities are fp32; g is rounded to bf16 before tanh(f + g) exactly as in rnnt_greedy_step.
 * rnnt_greedy_decode_workspace_bytes returns 0 when the shape is not covered (Hp % 8, H % 8, slices that do not fit the
 * shared memory of the co-resident CTAs); callers then use rnnt_greedy_step. */
size_t rnnt_greedy_decode_workspace_bytes(int B, int V, int H, int Hp);
int rnnt_greedy_decode_lstm(const void* f, const int32_t* lens, const void* W, const float* bias,
                            const float* gate_table, const void* W_hh, const void* W_proj, const float* bias_proj,
                            int B, int Tmax, int V, int H, int Hp, int blank, int max_symbols, int32_t* sym,
                            int sym_cap, int32_t* n_sym, void* workspace, size_t workspace_bytes, void* stream);

/* The same for a stack of 1..3 LSTM layers (torch.nn.LSTM(num_layers=n), unidirectional).  Layer 0 is described as above
 * (gate_table, W_hh); for the upper layers l = 1..n-1
 *   W_upper    bf16 [n-1][4*Hp][2*Hp]   [W_ih_l | W_hh_l] side by side (torch layouts)
 *   bias_upper f32  [n-1][4*Hp]         b_ih_l + b_hh_l
 * (both may be NULL when n_layers == 1); W_proj projects the top layer.  An utterance that emits blank keeps the state of
 * every layer. */
size_t rnnt_greedy_decode_stack_workspace_bytes(int B, int V, int H, int Hp, int n_layers);
int rnnt_greedy_decode_lstm_stack(const void* f, const int32_t* lens, const void* W, const float* bias,
                                  const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                                  const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax,
                                  int V, int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap,
                                  int32_t* n_sym, void* workspace, size_t workspace_bytes, void* stream);

/* The same loop for a GRU prediction network (torch.nn.GRU: r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise,
 * n = tanh(W_in x + b_in + r (W_hn h + b_hn)), h' = (1 - z) n + z h).  Arguments as for rnnt_greedy_decode_lstm_stack with
 * FOUR rows per hidden unit in the order (r, z, n_hidden, n_input):
 *   gate_table f32 [V+1][4*Hp]   (W_ir e_v + b_ir + b_hr | W_iz e_v + b_iz + b_hz | b_hn | W_in e_v + b_in)
 *   W_hh       bf16 [4*Hp][Hp]   (W_hr | W_hz | W_hn | 0)
 *   W_upper    bf16 [n-1][4*Hp][2*Hp]  rows ([W_ir|W_hr] | [W_iz|W_hz] | [0|W_hn] | [W_in|0]) of layer l
 *   bias_upper f32  [n-1][4*Hp]        (b_ir + b_hr | b_iz + b_hz | b_hn | b_in)
 * so that the data movement is the LSTM's; h is kept in fp32 for the blend and rounded to bf16 as a tensor-core operand. */
int rnnt_greedy_decode_gru_stack(const void* f, const int32_t* lens, const void* W, const float* bias,
                                 const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                                 const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax,
                                 int V, int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap,
                                 int32_t* n_sym, void* workspace, size_t workspace_bytes, void* stream);

/* Debug / test hooks (not part of the drop-in surface). */
int rnnt_debug_copy_stats(const void* workspace, int B, int Tmax, int Umax, int V, int H, float* lp_blank,
                          float* lp_label, float* c_blank, float* c_label, float* lnp_beta, void* stream);
void rnnt_debug_set(const char* key, int value);          /* "slab_tiles", "time_kernels", "reset_launches",
                                                              "path" (1 persistent kernels, 0 per-slab kernels),
                                                              "ring_slots" (2..4), "gemm_dbg" (bring-up switches),
                                                              "prune" (-1: the backward pass walks every tile; default 0: only
                                                              tiles with non-zero occupancy), "prune_log2_eps" (threshold 2^v),
                                                              "keep" (0: *_keep / *_kept ignore their buffer) */
long long rnnt_debug_get(const char* key);                /* "launches": kernels launched since the last reset */
/* In "time_kernels" mode every kernel launch is bracketed by CUDA events on its stream; this call
 * synchronises, sums the durations per kernel class (hgen, joint_fwd, joint_dz, joint_dh, joint_dw,
 * lattice, coefs, misc, joint_bwd_mega) into ms[0..9) / count[0..9) and clears the record (n >= 9). */
int rnnt_debug_kernel_times(double* ms, long long* count, int n);
/* Bring-up: %globaltimer stamps (8 per CTA) of the last tcgen05 GEMM launch made with gemm_dbg & 4. */
int rnnt_debug_read_prof(unsigned long long* out, int n);
/* After rnnt_fused_backward on `workspace`: out2[0] = lattice tiles the backward pass walked (those with non-zero arc
 * occupancy), out2[1] = tiles in the batch.  Synchronises the device. */
int rnnt_debug_read_active_tiles(const void* workspace, int B, int Tmax, int Umax, int V, int H, int* out2);
/* RNNT_PROFILE builds: per-CTA phase cycles of the mega-kernel's dh epilogue (8 values per CTA), optionally reset. */
int rnnt_debug_read_prof3(unsigned long long* out, int n, int reset);
/* Cluster decode with rnnt_debug_set("decode_prof", 1): SM cycles the epilogue of cluster 0 / rank 0 spent per stage,
 * summed over steps (out[0..9)), and the step count (out[9]). */
int rnnt_debug_decode_prof(unsigned long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* RNNT_B200_H_ */
