// Internal launch interface between the C-ABI (capi.cu) and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace rnnt {

// Function attributes (dynamic shared memory opt-in) and occupancy answers are per DEVICE, the library is per
// process: every cached flag / value is kept per device ordinal so that one process can drive several GPUs.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

// ---- lattice.cu -------------------------------------------------------------------------------
// alpha and beta wavefronts (one CTA per utterance and direction), writes alpha/beta (diagonal
// layout), loss[b] = -ln P(y|x) and lnp_beta[b] = beta[0,0] (consistency check).
// alpha / beta are fp64 (see lattice.cu, "Precision"); lnp64[b] = ln P(y|x) in fp64 for the coefficient kernel.
void launch_lattice_alpha_beta(const Lattice& L, const float* lpb, const float* lpl, double* alpha, double* beta,
                               float* loss, float* lnp_beta, double* lnp64, cudaStream_t s);
int lattice_max_columns();   // largest U + 1 the wavefront kernel covers
// c1 (blank-arc occupancy) and c2 (label-arc occupancy) per cell, diagonal layout.
void launch_lattice_coefs(const Lattice& L, const float* lpb, const float* lpl, const double* alpha,
                          const double* beta, const double* lnp64, float* c1, float* c2, cudaStream_t s);

void launch_nat_to_diag(const Lattice& L, const float* a_nat, const float* b_nat, float* a_diag, float* b_diag,
                        cudaStream_t s);
// Backward-pass tile list: flags[tile] = (max over the tile's cells of (c1 + c2) * |grad_loss[b]|) > eps, then the
// indices of the flagged tiles in increasing order (active) and their number (n_active[0]; n_active[1] = n_tiles).
void launch_tile_activity(const Lattice& L, const float* c1, const float* c2, const float* grad_loss, float eps, int* flags,
                          cudaStream_t s);
void launch_compact_tiles(const int* flags, int n_tiles, int* active, int* n_active, cudaStream_t s);

void launch_diag_to_nat(const Lattice& L, const float* a_diag, const float* b_diag, float* a_nat, float* b_nat,
                        cudaStream_t s);

// ---- joint.cu ---------------------------------------------------------------------------------
struct JointDims {
  int V, H;   // vocabulary (incl. blank), joint width
  int Vp;     // dz slab pitch (V rounded up to 64)
  int blank;
  int Umax;   // label tensor pitch
};

// h[row,:] = bf16(tanh(f[b,t,:] + g[b,u,:])) for every row of tiles [tile0, tile0 + n_tiles); padded rows = 0.
void launch_hgen(const Lattice& L, const __nv_bfloat16* f, const __nv_bfloat16* g, __nv_bfloat16* hslab, int tile0,
                 int n_tiles, int H, cudaStream_t s);

// Wt[h][v] = W[v][h] (pitch Vp, zero padded)
void launch_transpose_w(const __nv_bfloat16* W, __nv_bfloat16* Wt, int V, int H, int Vp, int perm, cudaStream_t s);

struct FwdArgs {
  const float* bias;   // [V] or null
  const int* y;        // [B][Umax]
  float* lse_tile;     // [n_tiles_total * 128]
  float* lpb;          // diagonal layout
  float* lpl;          // diagonal layout
};
// logits tile = hslab . W^T (tcgen05), fused online log-softmax; keeps only lse, lp_blank, lp_label.
void launch_joint_fwd(const Lattice& L, const JointDims& d, const CUtensorMap& tm_h, const CUtensorMap& tm_w,
                      const FwdArgs& a, int tile0, int n_tiles, int nc, cudaStream_t s);

struct DzArgs {
  const float* bias;
  const int* y;
  const float* lse_tile;
  const float* lpb;
  const float* lpl;
  const float* c1;
  const float* c2;
  const float* grad_loss;  // [B]
  float* db;               // [V] accumulated with red.add
};
// Recomputes the logits tile and writes dz = dL/dlogits (bf16) into the dz slab; accumulates db.
void launch_joint_dz(const Lattice& L, const JointDims& d, const CUtensorMap& tm_h, const CUtensorMap& tm_w,
                     const CUtensorMap& tm_dz_store, const DzArgs& a, int tile0, int n_tiles, int nc,
                     cudaStream_t s);

struct DhArgs {
  const __nv_bfloat16* hslab;
  float* df;  // [B][Tmax][H]
  float* dg;  // [B][U1max][H]
};
// dh = dz . W (tcgen05, A = dz slab, B = W^T), dpre = dh * (1 - h^2), tile-reduced into df / dg.
void launch_joint_dh(const Lattice& L, const JointDims& d, const CUtensorMap& tm_dz, const CUtensorMap& tm_wt,
                     const DhArgs& a, int tile0, int n_tiles, int nc, cudaStream_t s);

// dW += dz^T . h over the slab rows (tcgen05, both operands MN-major), split-K across CTAs, red.add into dW.
void launch_joint_dw(const JointDims& d, const CUtensorMap& tm_dz_mn, const CUtensorMap& tm_h_mn, float* dW,
                     int n_tiles, int n_ctas, cudaStream_t s);

// Greedy decode step: for each active utterance, k = argmax_v W . tanh(f[b,t_b,:] + g[b,:]) + bias.
void launch_greedy_argmax(const __nv_bfloat16* f, const __nv_bfloat16* g, const __nv_bfloat16* W, const float* bias,
                          const int* t_idx, int* out_k, int B, int Tmax, int V, int H, cudaStream_t s);

// ---- persist.cu -------------------------------------------------------------------------------
constexpr int kMaxPersistCtas = 148;  // scratch is sized for this many CTAs (one per B200 SM)
constexpr int kMaxRingSlots = 4;

struct FwdPArgs {
  Lattice L;
  int dbg;
  int csize;                  // CTAs per cluster: 2 (one CTA pair) or 4 (two pairs, W multicast)
  int hgen_warps;             // 4, or 8 for narrow vocabularies (the pass is bound by the tanh evaluations)
  int keep_z;                 // != 0: the base-2 logits are also stored as fp16 through the tm_z map (<= 4096 columns)
  __nv_bfloat16* hkeep;       // keep_z: [tiles][128][H] h tiles are written here (and read back as the A operand through
                              // tm_hscratch, which then maps this buffer) instead of the per-CTA scratch
  int n_tiles_total;
  int V, H;
  int nc, n_chunks, k_blocks;
  int blank, Umax;
  const __nv_bfloat16* f;     // [B][Tmax][H]
  const __nv_bfloat16* g;     // [B][U1max][H]
  __nv_bfloat16* hscratch;    // [n_ctas][2][128][H]  per-CTA double-buffered h tiles (L2-resident)
  const float* bias;
  const int* y;
  float* lse_tile;
  float* lpb;
  float* lpl;
};
// One persistent launch for the whole batch: hgen + logits (tcgen05) + online log-softmax.
void launch_fwd_persist(const CUtensorMap& tm_hscratch, const CUtensorMap& tm_w, const CUtensorMap& tm_z, const FwdPArgs& a,
                        int n_ctas, cudaStream_t s);
int smem_bytes_fwd_persist();
int max_ctas_fwd_persist(int csize, int hgen_warps = 4);
int read_persist_prof(unsigned long long* out, int n);
int read_persist_prof3(unsigned long long* out, int n, int reset);
int get_gemm_dbg();

struct BwdPArgs {
  Lattice L;
  int dbg;
  int csize;                   // CTAs per cluster (2 or 4)
  int cons_share;              // 4-clusters of consumers share their h boxes (n_vt even)
  int n_tiles_total;
  const int* active_tiles;     // compacted list of the tiles that carry occupancy (device), or null = all tiles
  const int* n_active;         // its length (device)
  int V, H, Vp;
  int nc_v, n_chunks_v, kb_h;  // dz pass: N chunks over V, k-blocks over H
  int nc_h, n_chunks_h, kb_v;  // dh pass: N chunks over H, k-blocks over Vp
  int blank, Umax;
  int P, C, KG, NS;            // producer pairs, consumer pairs, K-groups, ring slots per producer pair
  int n_vt, n_ht, n_out;       // dW blocks of 256 (V) x 512 (H); n_out = n_vt * n_ht consumers per K-group
  const __nv_bfloat16* f;
  const __nv_bfloat16* g;
  __nv_bfloat16* h_ring;       // [P][NS][256][H]
  __nv_bfloat16* dz_ring;      // [P][NS][256][Vp]
  const float* bias;
  const int* y;
  const float* lse_tile;
  const float* lpb;
  const float* lpl;
  const float* c1;
  const float* c2;
  const float* grad_loss;
  float* db;
  float* df;
  float* dg;
  float* dW;
  const uint16_t* zlog;        // fp16 base-2 logits [tiles][128][Vp] kept by the forward pass, or null (recompute them)
  const __nv_bfloat16* hkeep;  // with zlog: h [tiles][128][H] kept by the forward pass (tm_h_mn maps it; no hgen, no h ring)
  int zcols;                   // columns of a zlog row the forward pass wrote (multiple of 32, <= Vp)
  unsigned* ready;             // [P][NS] per use: += 1 per producer epilogue warp (2 CTAs x 8), or per CTA (2) with zlog
  unsigned* done;              // [P][NS] += 1 per consumer pair per use
};
// One persistent launch for the whole backward pass (see persist.cu).
void launch_bwd_mega(const CUtensorMap& tm_h, const CUtensorMap& tm_w, const CUtensorMap& tm_dz, const CUtensorMap& tm_wt,
                     const CUtensorMap& tm_dz_mn, const CUtensorMap& tm_h_mn, const CUtensorMap& tm_dz_st,
                     const CUtensorMap& tm_zl, const BwdPArgs& a, int n_ctas, cudaStream_t s);
int smem_bytes_bwd_mega();
int max_ctas_bwd_mega(int csize);
int bwd_mega_cooperative();   // 1 while the cooperative (co-scheduled) launch is in use
void set_bwd_mega_cooperative(int v);

// One greedy decode step with its bookkeeping (see joint.cu::greedy_step_kernel).
void launch_greedy_step(const __nv_bfloat16* f, const float* g, const __nv_bfloat16* W, const float* bias, const int* lens,
                        int* t_cur, int* emitted, int* n_sym, int* sym, int sym_cap, int* is_sym, int* label, int* active,
                        int B, int Tmax, int V, int H, int blank, int max_symbols, cudaStream_t s);


// ---- decode.cu --------------------------------------------------------------------------------
// Whole greedy decode loop (LSTM prediction cell + projection + joint argmax + bookkeeping) in one cooperative launch.
struct DecodeArgs {
  int B, Bp, Tmax, V, H, Hp, blank, S, sym_cap, max_steps;
  int nJ, nslJ, kbJ;       // joint:      vocabulary rows per CTA, CTAs with a slice, k-blocks over H
  int nu, nL, nslL, kbL;   // LSTM cell:  hidden units per CTA, nL = 4 nu gate rows, slices, k-blocks over Hp
  int nP, nslP, kbP;       // projection: output columns per CTA, slices, k-blocks over Hp
  int n_stages;            // activation ring depth (16 KB stages)
  int o_wj, o_wl, o_wp, o_c, o_h, o_g, o_state, o_bars;  // shared-memory offsets from the 1024-aligned base
  const __nv_bfloat16* f;  // [B][Tmax][H]
  const int* lens;         // [B] device
  const float* bias_j;     // [V] or null
  const float* table;      // [V+1][4 Hp]  W_ih . emb[v] + b_ih + b_hh, torch gate order i f g o; row V = start of sequence
  const float* bias_p;     // [H] or null
  __nv_bfloat16* hj;       // [Bp][H]      bf16(tanh(f[b, t_b] + g[b]))
  __nv_bfloat16* hbuf;     // [2][Bp][Hp]  hidden state, double buffered
  unsigned long long* amax;  // [2][Bp]    packed (ordered logit, ~index) argmax keys
  unsigned* gbar;          // grid barrier counter
  int* sym;                // [B][sym_cap]
  int* n_sym;              // [B]
};
void launch_permute_whh(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int nu, int n_rows, cudaStream_t s);
int max_ctas_greedy_decode(int smem_bytes);
cudaError_t launch_greedy_decode(const CUtensorMap& tm_hj, const CUtensorMap& tm_hbuf, const CUtensorMap& tm_wj,
                                 const CUtensorMap& tm_wl, const CUtensorMap& tm_wp, const DecodeArgs& a, int n_ctas,
                                 int smem_bytes, cudaStream_t s);
void set_decode_cooperative(int v);

// Cluster variant: one thread-block cluster per 16 utterances, activations resident, weights streamed (see decode.cu).
struct ClusterDecodeArgs {
  int B, Tmax, V, H, Hp, blank, S, sym_cap, max_steps;
  int C;                   // CTAs per cluster
  int RJ, RP, up;          // per-CTA vocabulary rows, projection rows, hidden units (multiples of 64)
  int mtJ, mtP, mtL;       // 128-row accumulator tiles per CTA and product
  int kbH, kbHp;           // k-blocks over H / Hp
  int n_stages, tmem_cols;
  int NL;                  // recurrent layers of the prediction network (1..3)
  int cell;                // 0 = LSTM (gate rows i, f, g, o), 1 = GRU (gate rows r, z, n_hidden, n_input)
  int o_hj, o_layers, layer_stride, o_gates, o_amax, o_part, o_state, o_bars;
  const __nv_bfloat16* f;
  const int* lens;
  const float* bias_j;
  const float* table;
  const float* bias_up;    // [NL-1][4 Hp]  b_ih + b_hh of the upper layers
  const float* bias_p;
  int* sym;
  int* n_sym;
  int res_p;               // leading k-blocks of this CTA's W_proj tile resident in TMEM columns [res_col, + 32 res_p)
  int res_col;
  int res_j, res_j_col;    // leading k-blocks of the first vocabulary tile resident in TMEM columns [res_j_col, + 32 res_j)
  const __nv_bfloat16* w_proj;    // [H][Hp] (only read when res_p)
  const __nv_bfloat16* w_joint;   // [V][H]  (only read when res_j)
  int l_late;              // experiment: issue all of W_hh . h(s+1) after the vocabulary product instead of tile 0 before it
  int prof;                // != 0: cluster 0 / rank 0 sums clock64 cycles per epilogue stage into g_dec_prof
};
int read_decode_prof(unsigned long long* out, int n);
void launch_permute_whh_cluster(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int up, int n_rows, cudaStream_t s);
int max_clusters_greedy_decode(int smem_bytes, int C);
void launch_permute_wup_cluster(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int up, int C, int kb, int n_rows,
                                cudaStream_t s);
cudaError_t launch_greedy_decode_cluster(const CUtensorMap& tm_wj, const CUtensorMap& tm_wl, const CUtensorMap& tm_wu,
                                         const CUtensorMap& tm_wp, const ClusterDecodeArgs& a, int n_clusters, int smem_bytes,
                                         cudaStream_t s);


void set_gemm_dbg(int v);
int read_gemm_prof(unsigned long long* out, int n);
int smem_bytes_fwd(int nc_total);
int smem_bytes_dz(int nc_total);
int smem_bytes_dh();
int smem_bytes_dw();

}  // namespace rnnt
