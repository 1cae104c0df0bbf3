// C ABI (include/rnnt_b200.h): argument validation, workspace carving, TMA tensor maps and the
// per-slab kernel schedule.  No torch types, no allocation, no host synchronisation.
//
// Schedule.  Lattice rows are processed in slabs of `slab_tiles` tiles (one tile per SM).  Per slab
//   forward : hgen -> joint_fwd                       (h slab stays L2-resident between the two)
//   backward: hgen -> joint_dz -> joint_dh -> joint_dw (h and dz slabs are L2-resident ring buffers;
//             the B*T*U*V logits / dlogits tensors never exist in HBM)
// followed / preceded by the alpha-beta lattice kernels over the whole batch.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/rnnt_b200.h"
#include "launch.h"

namespace {

using namespace rnnt;

thread_local char g_err[512] = "";
int g_slab_tiles_override = 0;
int g_path = 1;  // 1 = persistent kernels (persist.cu), 0 = per-slab kernels (joint.cu)
int g_ring_slots = 2;
int g_kgk_override = 0;  // the same for the kept-logits schedule
int g_kg_override = 0;   // bring-up: K-groups of dW consumers in the backward mega-kernel (0 = plan's choice)
int g_cluster = 2;      // forward kernel: CTAs per cluster; 4 = two CTA pairs sharing W through TMA multicast
                        // (measured slower: only 132 of the 148 SMs can host 4-clusters)
int g_keep = 1;      // 0: a buffer passed to the *_keep / *_kept calls is ignored (recompute schedule)
int g_prune = 0;            // backward mega-kernel: >= 0 walks only tiles that carry occupancy; -1 = every tile
int g_prune_log2_eps = -100000;  // occupancy threshold 2^v (default: exact zero only)
int g_fwd_hgen_warps = 0;  // 0 = by shape; 4 / 8 force the forward kernel's number of hgen warps
int g_cluster_bwd = 2;  // backward mega-kernel: 4-clusters cannot all be co-resident (132 of 148 SMs), so pairs

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) return fail(RNNT_ERR_CUDA, "%s -> %s", #x, cudaGetErrorString(e_)); \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- launch accounting (rnnt_debug_get / rnnt_debug_kernel_times) ---------------------------------
enum KClass { K_HGEN = 0, K_FWD, K_DZ, K_DH, K_DW, K_LATTICE, K_COEFS, K_MISC, K_BWD_MEGA, K_NCLASS };
long long g_launches[K_NCLASS] = {0};
bool g_time_kernels = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_pairs[K_NCLASS];
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t pool_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// Wraps one kernel launch: counts it and, in timing mode, brackets it with events on the launch stream.
#define KLAUNCH(cls, stream, call)                                    \
  do {                                                                \
    ++g_launches[cls];                                                \
    if (g_time_kernels) {                                             \
      cudaEvent_t e0_ = pool_event(), e1_ = pool_event();             \
      cudaEventRecord(e0_, stream);                                   \
      call;                                                           \
      cudaEventRecord(e1_, stream);                                   \
      g_pairs[cls].push_back({e0_, e1_});                             \
    } else {                                                          \
      call;                                                           \
    }                                                                 \
  } while (0)

int sm_count() {
  static int cache[kMaxDevices] = {};
  const int dev = current_device();
  int& n = cache[dev];
  if (n == 0) {
    int v = 0;
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
  }
  return n;
}

struct Plan {
  int B, Tmax, Umax, U1, V, H, Vp, D;
  int max_tiles, slab_tiles;
  size_t o_prefix, o_flens, o_ylens, o_lse, o_lpb, o_lpl, o_alpha, o_beta, o_c1, o_c2, o_lnpb, o_lnp64, o_wt, o_h, o_dz, o_hs, o_hring, o_dzring, o_flags, o_tflags, o_active, o_nactive;
  int mega_ok, n_vt, n_ht, n_out, KG, C, P, NS;
  int keep_ok, KGk, Ck, Pk;      // role split of the mega-kernel when the forward pass kept the logits (no recompute GEMM)
  size_t state_bytes, total, kept_bytes, o_keep_h;
};

Plan make_plan(int B, int Tmax, int Umax, int V, int H) {
  Plan p{};
  p.B = B; p.Tmax = Tmax; p.Umax = Umax; p.U1 = Umax + 1; p.V = V; p.H = H;
  p.Vp = static_cast<int>(align_up(V, 64));
  p.D = Tmax + p.U1;
  p.max_tiles = B * ((Tmax + kTT - 1) / kTT) * ((p.U1 + kTU - 1) / kTU);
  p.slab_tiles = g_slab_tiles_override > 0 ? g_slab_tiles_override : 148;
  if (p.slab_tiles > p.max_tiles) p.slab_tiles = p.max_tiles;
  p.slab_tiles = (p.slab_tiles + 1) / 2 * 2;  // CTA pairs: a slab holds an even number of tiles
  const size_t cells = static_cast<size_t>(B) * p.D * p.U1;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  p.o_prefix = take(sizeof(int) * (B + 1));
  p.o_flens = take(sizeof(int) * B);
  p.o_ylens = take(sizeof(int) * B);
  p.o_lse = take(sizeof(float) * static_cast<size_t>(p.max_tiles) * kTileRows);
  p.o_lpb = take(sizeof(float) * cells);
  p.o_lpl = take(sizeof(float) * cells);
  p.o_c1 = take(sizeof(float) * cells);
  p.o_c2 = take(sizeof(float) * cells);
  // Everything up to here is the STATE the forward call leaves for the backward call (rnnt_fused_state_bytes); all
  // that follows is scratch that either call may overwrite.
  p.state_bytes = o;
  p.o_alpha = take(sizeof(double) * cells);
  p.o_beta = take(sizeof(double) * cells);
  p.o_lnpb = take(sizeof(float) * B);
  p.o_lnp64 = take(sizeof(double) * B);
  p.o_wt = take(2 * align_up(H, 32) * p.Vp);       // W^T, rows padded to whole groups of 32 (joint.cu::transpose_w_kernel)
  p.o_h = take(2 * static_cast<size_t>(p.slab_tiles) * kTileRows * H);
  p.o_dz = take(2 * static_cast<size_t>(p.slab_tiles) * kTileRows * p.Vp);
  p.o_hs = take(2 * static_cast<size_t>(kMaxPersistCtas) * 2 * kTileRows * H);
  // backward mega-kernel: role split of the 74 CTA pairs (see persist.cu)
  p.n_vt = (p.Vp + 255) / 256; p.n_ht = (H + 511) / 512; p.n_out = p.n_vt * p.n_ht;
  // Producer : consumer split of the 74 CTA pairs from a cycle model of one 256-row pair-tile (all MMAs are M = 256
  // pair MMAs of ~128 cycles; epilogue costs per 256-column chunk measured with scripts/prof_mega_chunks.py).  A
  // producer pair runs the dz and dh passes, MMAs overlapped with the epilogue of the previous chunk; the consumers of
  // one K-group together add the pair-tile to every dW block.  V = H = 1024: 74 k vs 33 k cycles -> 24 consumer pairs
  // (3 K-groups of 8 blocks); V = 29, H = 512 (epilogue-bound producers): 23 k vs 4 k -> 11 (measured optimum 11-13,
  // scripts/kg_sweep.py; the flop-balanced 25 : 49 split it replaces ran configs[1] 14 % slower).
  {
    const double kb_h = (H + 63) / 64, kb_v = p.Vp / 64;
    const double ch_v = (p.Vp + 255) / 256, ch_h = (H + 255) / 256;
    const double mma_p = (ch_v * kb_h + ch_h * kb_v) * 4 * 128;
    const double epi_p = 8000.0 * p.Vp / 256 + 8800.0 * H / 256;
    const double prod = mma_p > epi_p ? mma_p : epi_p;
    double cons = 0;   // per block: 4 k-blocks x (1 or 2 accumulators) x 4 MMAs
    for (int hb = 0; hb < p.n_ht; ++hb) cons += p.n_vt * 4 * ((H - hb * 512 > 256) ? 8 : 4) * 128.0;
    const double c_ideal = (kMaxPersistCtas / 2) * cons / (cons + prod);
    p.KG = static_cast<int>(c_ideal / p.n_out + 0.5);
    // kept logits: the producers run the dh pass only (dz is a streaming pass of the front-end warps)
    const double mma_k = ch_h * kb_v * 4 * 128, epi_k = 8800.0 * H / 256;
    const double prod_k = mma_k > epi_k ? mma_k : epi_k;
    p.KGk = static_cast<int>((kMaxPersistCtas / 2) * cons / (cons + prod_k) / p.n_out + 0.5);
  }
  if (g_kg_override > 0) p.KG = g_kg_override;
  if (g_kgk_override > 0) p.KGk = g_kgk_override;
  if (p.KG < 1) p.KG = 1;
  if (p.KGk < 1) p.KGk = 1;
  p.C = p.n_out * p.KG;
  p.P = kMaxPersistCtas / 2 - p.C;
  p.Ck = p.n_out * p.KGk;
  p.Pk = kMaxPersistCtas / 2 - p.Ck;
  if (p.Pk > p.P) { p.Pk = p.P; p.KGk = p.KG; p.Ck = p.C; }   // the rings are sized for P producer pairs
  p.NS = g_ring_slots;
  p.mega_ok = p.C <= 40 && p.P >= 1 && p.Vp <= 4096;   // 16 V chunks / 4096 bias columns in the mega-kernel
  p.keep_ok = p.mega_ok && p.Ck <= 40 && p.Pk >= 1;
  // fp16 base-2 logits and bf16 h of every lattice row, kept from the forward to the backward pass in a buffer of the caller
  p.o_keep_h = align_up(2 * static_cast<size_t>(p.max_tiles) * kTileRows * p.Vp, 1024);
  p.kept_bytes = p.Vp <= 4096 ? p.o_keep_h + 2 * static_cast<size_t>(p.max_tiles) * kTileRows * H : 0;
  p.o_tflags = take(sizeof(int) * static_cast<size_t>(p.max_tiles));   // backward-pass tile list (lattice.cu)
  p.o_active = take(sizeof(int) * static_cast<size_t>(p.max_tiles));
  p.o_nactive = take(sizeof(int) * 4);
  p.o_hring = p.o_dzring = p.o_flags = 0;
  if (p.mega_ok) {
    p.o_hring = take(2 * static_cast<size_t>(p.P) * kMaxRingSlots * 2 * kTileRows * H);
    p.o_dzring = take(2 * static_cast<size_t>(p.P) * kMaxRingSlots * 2 * kTileRows * p.Vp);
    p.o_flags = take(sizeof(unsigned) * 2 * (kMaxPersistCtas / 2) * kMaxRingSlots);
  }
  p.total = o;
  return p;
}

int check_dims(int B, int Tmax, int Umax, int V, int H) {
  if (B < 1 || Tmax < 1 || Umax < 0) return fail(RNNT_ERR_INVALID_ARGUMENT, "B=%d Tmax=%d Umax=%d out of range", B, Tmax, Umax);
  if (V < 1 || V > 8192) return fail(RNNT_ERR_UNSUPPORTED, "V=%d must be in [1, 8192]", V);
  if (H < 8 || H % 8 != 0) return fail(RNNT_ERR_UNSUPPORTED, "H=%d must be a positive multiple of 8", H);
  if (Umax + 1 > lattice_max_columns())
    return fail(RNNT_ERR_UNSUPPORTED, "Umax+1=%d must be <= %d", Umax + 1, lattice_max_columns());
  return RNNT_OK;
}

int check_lens(const int32_t* fl, const int32_t* yl, int B, int Tmax, int Umax) {
  if (!fl || !yl) return fail(RNNT_ERR_INVALID_ARGUMENT, "length arrays must not be NULL");
  for (int b = 0; b < B; ++b) {
    if (fl[b] < 1 || fl[b] > Tmax) return fail(RNNT_ERR_INVALID_ARGUMENT, "f_lens[%d]=%d must be in [1, %d]", b, fl[b], Tmax);
    if (yl[b] < 0 || yl[b] > Umax) return fail(RNNT_ERR_INVALID_ARGUMENT, "y_lens[%d]=%d must be in [0, %d]", b, yl[b], Umax);
  }
  return RNNT_OK;
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D bf16 row-major tensor [rows][cols] with `pitch` elements per row; box = box_cols x box_rows, 128B swizzle.
int make_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_cols,
             uint32_t box_rows) {
  auto enc = get_encode();
  if (!enc) return fail(RNNT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(RNNT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu pitch=%llu box=%ux%u", (int)r,
                (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch, box_cols, box_rows);
  return RNNT_OK;
}

// fp16 [rows][cols] map for the kept logits: 32 x 32 boxes (64-byte rows), 64-byte swizzle
// (forward pass, stores), or 64 x 128 boxes with the 128-byte swizzle (backward pass, loads)
int make_map_f16(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, bool load_boxes) {
  auto enc = get_encode();
  if (!enc) return fail(RNNT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {load_boxes ? 64u : 32u, load_boxes ? 128u : 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, load_boxes ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(RNNT_ERR_CUDA, "cuTensorMapEncodeTiled (kept logits) failed (%d) cols=%llu rows=%llu", (int)r,
                (unsigned long long)cols, (unsigned long long)rows);
  return RNNT_OK;
}

int chunk_cols(int n) {
  int nc = (n + 31) / 32 * 32;
  return nc > 256 ? 256 : nc;
}

struct Ws {
  uint8_t* base;
  const Plan& p;
  template <class T> T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
};

Lattice make_lattice(const Ws& w, int n_tiles_total) {
  Lattice L{};
  L.tile_prefix = w.at<int>(w.p.o_prefix);
  L.f_lens = w.at<int>(w.p.o_flens);
  L.y_lens = w.at<int>(w.p.o_ylens);
  L.B = w.p.B; L.Tmax = w.p.Tmax; L.U1max = w.p.U1; L.D = w.p.D; L.n_tiles_total = n_tiles_total;
  return L;
}

int upload_lengths(const Ws& w, const int32_t* fl, const int32_t* yl, int* n_tiles_out, cudaStream_t s) {
  const int B = w.p.B;
  std::vector<int> prefix(B + 1, 0);
  for (int b = 0; b < B; ++b)
    prefix[b + 1] = prefix[b] + ((fl[b] + kTT - 1) / kTT) * ((yl[b] + 1 + kTU - 1) / kTU);
  *n_tiles_out = prefix[B];
  // pageable sources: the runtime stages them before returning, so the host buffers may die after the call
  CUDA_TRY(cudaMemcpyAsync(w.at<int>(w.p.o_prefix), prefix.data(), sizeof(int) * (B + 1), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(w.at<int>(w.p.o_flens), fl, sizeof(int) * B, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(w.at<int>(w.p.o_ylens), yl, sizeof(int) * B, cudaMemcpyHostToDevice, s));
  return RNNT_OK;
}

int count_tiles(const int32_t* fl, const int32_t* yl, int B) {
  int n = 0;
  for (int b = 0; b < B; ++b) n += ((fl[b] + kTT - 1) / kTT) * ((yl[b] + 1 + kTU - 1) / kTU);
  return n;
}

// ---- greedy decode (decode.cu): N-split of the three per-step GEMMs over the co-resident CTAs ----------------------
struct DecPlan {
  int ok;
  int Bp, nJ, nslJ, kbJ, nu, nL, nslL, kbL, nP, nslP, kbP, n_stages;
  int o_wj, o_wl, o_wp, o_c, o_h, o_g, o_state, o_bars, smem;
  size_t w_gbar, w_amax, w_hj, w_hbuf, w_whh, w_total;
};

int slice16(int n_total, int G) {  // smallest multiple of 16 that covers n_total with at most G slices
  int n = 16;
  while ((n_total + n - 1) / n > G) n += 16;
  return n;
}

DecPlan make_dec_plan(int B, int V, int H, int Hp) {
  DecPlan d{};
  if (B < 1 || B > 2048 || V < 1 || H < 8 || H % 8 || Hp < 8 || Hp % 8) return d;
  const int G = sm_count();
  d.Bp = static_cast<int>(align_up(B, 128));
  d.nJ = slice16(V, G); d.nslJ = (V + d.nJ - 1) / d.nJ; d.kbJ = (H + 63) / 64;
  d.nP = slice16(H, G); d.nslP = (H + d.nP - 1) / d.nP; d.kbP = (Hp + 63) / 64;
  d.nu = 4;
  while ((Hp + d.nu - 1) / d.nu > G) d.nu += 4;
  d.nL = 4 * d.nu; d.nslL = (Hp + d.nu - 1) / d.nu; d.kbL = (Hp + 63) / 64;
  if (d.nJ > 256 || d.nP > 256 || d.nL > 256) return d;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return static_cast<int>(r); };
  d.o_wj = take(static_cast<size_t>(d.kbJ) * d.nJ * 128);
  d.o_wl = take(static_cast<size_t>(d.kbL) * d.nL * 128);
  d.o_wp = take(static_cast<size_t>(d.kbP) * d.nP * 128);
  d.o_c = take(sizeof(float) * d.Bp * d.nu);
  d.o_h = take(2 * static_cast<size_t>(d.Bp) * d.nu);
  d.o_g = take(sizeof(float) * d.Bp * d.nP);
  d.o_state = take(sizeof(int) * 5 * B);
  d.o_bars = take(256);
  const size_t fixed = o + 1024;  // + alignment slack
  const size_t cap = 227 * 1024;
  if (fixed + 2 * 16384 > cap) return d;
  d.n_stages = static_cast<int>((cap - fixed) / 16384);
  if (d.n_stages > 8) d.n_stages = 8;
  const int ring = d.n_stages * 16384;  // the ring sits first (1024-aligned stages)
  d.o_wj += ring; d.o_wl += ring; d.o_wp += ring; d.o_c += ring; d.o_h += ring; d.o_g += ring; d.o_state += ring; d.o_bars += ring;
  d.smem = static_cast<int>(fixed) + ring;
  size_t w = 0;
  auto wtake = [&](size_t bytes) { size_t r = w; w = align_up(w + bytes, 1024); return r; };
  d.w_gbar = wtake(1024);
  d.w_amax = wtake(sizeof(unsigned long long) * 2 * d.Bp);
  d.w_hj = wtake(2 * static_cast<size_t>(d.Bp) * H);
  d.w_hbuf = wtake(2 * static_cast<size_t>(2) * d.Bp * Hp);
  d.w_whh = wtake(2 * static_cast<size_t>(d.nslL) * d.nL * Hp);
  d.w_total = w;
  d.ok = 1;
  return d;
}

// Cluster variant of the decode (decode.cu): rows of each weight matrix split over the C CTAs of a cluster.
struct CDecPlan {
  int ok;
  int C, NL, RJ, RP, up, mtJ, mtP, mtL, kbH, kbHp, n_stages, tmem_cols;
  int o_hj, o_layers, layer_stride, o_gates, o_amax, o_part, o_state, o_bars, smem;
  size_t w_whh, w_wup, w_total;
};

CDecPlan make_cdec_plan(int B, int V, int H, int Hp, int C, int NL = 1) {
  CDecPlan d{};
  if (B < 1 || V < 1 || H < 8 || H % 8 || Hp < 8 || Hp % 8 || C < 1 || C > 16 || NL < 1 || NL > 3) return d;
  d.NL = NL;
  auto up64 = [](int x) { return (x + 63) / 64 * 64; };
  d.C = C;
  d.up = up64((Hp + C - 1) / C); d.RP = up64((H + C - 1) / C); d.RJ = up64((V + C - 1) / C);
  d.mtL = d.up / 32; d.mtP = (d.RP + 127) / 128; d.mtJ = (d.RJ + 127) / 128;
  const int n_tiles = NL * d.mtL + d.mtP + d.mtJ;
  if (n_tiles > 10 || d.mtL > 4) return d;   // kMaxTiles / kMaxLTiles in decode.cu
  d.kbH = (H + 63) / 64; d.kbHp = (Hp + 63) / 64;
  d.tmem_cols = 32;
  while (d.tmem_cols < 32 * n_tiles) d.tmem_cols *= 2;   // two partial accumulators x 16 utterances per tile
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return static_cast<int>(r); };
  d.o_hj = take(static_cast<size_t>(C) * d.RP / 64 * 2048);
  // per layer: two h operand buffers, the cell state and this CTA's own units of h (all multiples of 1 KB: up % 64 == 0)
  d.layer_stride = static_cast<int>(2 * (static_cast<size_t>(C) * d.up / 64 * 2048) + static_cast<size_t>(d.up) * 16 * 4 +
                                    static_cast<size_t>(d.up) * 16 * 2);
  d.o_layers = take(static_cast<size_t>(NL) * d.layer_stride);
  d.o_gates = take(static_cast<size_t>(d.mtL) * 4 * 32 * 16 * sizeof(float));
  d.o_amax = take(static_cast<size_t>(C) * 16 * 8);
  d.o_part = take(4 * 16 * 8);
  d.o_state = take(512);
  d.o_bars = take(512);
  const size_t fixed = o + 1024, cap = 227 * 1024;
  if (fixed + 3 * 16384 > cap) return d;
  d.n_stages = static_cast<int>((cap - fixed) / 16384);
  if (d.n_stages > 12) d.n_stages = 12;
  const int ring = d.n_stages * 16384;
  d.o_hj += ring; d.o_layers += ring; d.o_gates += ring; d.o_amax += ring; d.o_part += ring; d.o_state += ring; d.o_bars += ring;
  d.smem = static_cast<int>(fixed) + ring;
  d.w_whh = 0;
  d.w_wup = align_up(2 * static_cast<size_t>(C) * 4 * d.up * Hp, 1024);
  d.w_total = d.w_wup + align_up(2 * static_cast<size_t>(NL - 1) * C * 4 * d.up * 2 * d.kbHp * 64, 1024);
  d.ok = 1;
  return d;
}

int g_decode_variant = 1;  // 1: one cluster per 16 utterances (default); 0: N-split over the grid with grid barriers
int g_decode_cluster = 8;
int g_decode_prof = 0;
int g_decode_l_late = 0;
int g_decode_res = 1;    // keep the projection weights resident in TMEM when they fit

}  // namespace

extern "C" {

int rnnt_abi_version(void) { return 3; }

const char* rnnt_last_error(void) { return g_err; }

void rnnt_debug_set(const char* key, int value) {
  if (!strcmp(key, "slab_tiles")) g_slab_tiles_override = value;
  if (!strcmp(key, "time_kernels")) g_time_kernels = value != 0;
  if (!strcmp(key, "gemm_dbg")) set_gemm_dbg(value);
  if (!strcmp(key, "path")) g_path = value;
  if (!strcmp(key, "decode_cooperative")) set_decode_cooperative(value);
  if (!strcmp(key, "decode_prof")) g_decode_prof = value;
  if (!strcmp(key, "decode_resident")) g_decode_res = value;
  if (!strcmp(key, "decode_l_late")) g_decode_l_late = value;
  if (!strcmp(key, "mega_kg")) g_kg_override = value;
  if (!strcmp(key, "mega_kg_kept")) g_kgk_override = value;
  if (!strcmp(key, "decode_variant") && (value == 0 || value == 1)) g_decode_variant = value;
  if (!strcmp(key, "decode_cluster") && value >= 1 && value <= 16) g_decode_cluster = value;
  if (!strcmp(key, "mega_cooperative")) set_bwd_mega_cooperative(value);  // 0: plain launch (ncu cannot replay cooperative launches)
  if (!strcmp(key, "cluster") && (value == 2 || value == 4)) g_cluster = value;
  if (!strcmp(key, "cluster_bwd") && (value == 2 || value == 4)) g_cluster_bwd = value;
  if (!strcmp(key, "fwd_hgen_warps")) g_fwd_hgen_warps = value;
  if (!strcmp(key, "prune")) g_prune = value;
  if (!strcmp(key, "keep")) g_keep = value;
  if (!strcmp(key, "prune_log2_eps")) g_prune_log2_eps = value;
  if (!strcmp(key, "ring_slots") && value >= 2 && value <= 4) g_ring_slots = value;
  if (!strcmp(key, "reset_launches")) for (int i = 0; i < K_NCLASS; ++i) g_launches[i] = 0;
}

long long rnnt_debug_get(const char* key) {
  if (!strcmp(key, "launches")) {
    long long n = 0;
    for (int i = 0; i < K_NCLASS; ++i) n += g_launches[i];
    return n;
  }
  if (!strcmp(key, "n_classes")) return K_NCLASS;
  if (!strcmp(key, "mega_cooperative")) return bwd_mega_cooperative();
  if (!strcmp(key, "max_ctas_fwd_c2")) return max_ctas_fwd_persist(2);
  if (!strcmp(key, "max_ctas_fwd_c4")) return max_ctas_fwd_persist(4);
  if (!strcmp(key, "max_ctas_mega_c2")) return max_ctas_bwd_mega(2);
  if (!strcmp(key, "max_ctas_mega_c4")) return max_ctas_bwd_mega(4);
  if (!strcmp(key, "decode_max_clusters_8")) return max_clusters_greedy_decode(216 * 1024, 8);
  if (!strcmp(key, "decode_max_clusters_16")) return max_clusters_greedy_decode(216 * 1024, 16);
  return -1;
}

int rnnt_debug_decode_prof(unsigned long long* out, int n) { return read_decode_prof(out, n); }

int rnnt_debug_read_active_tiles(const void* workspace, int B, int Tmax, int Umax, int V, int H, int* out2) {
  if (check_dims(B, Tmax, Umax, V, H) != RNNT_OK) return RNNT_ERR_INVALID_ARGUMENT;
  const Plan p = make_plan(B, Tmax, Umax, V, H);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out2, static_cast<const uint8_t*>(workspace) + p.o_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost));
  return RNNT_OK;
}

int rnnt_debug_read_prof3(unsigned long long* out, int n, int reset) { return read_persist_prof3(out, n, reset); }

int rnnt_debug_read_prof(unsigned long long* out, int n) { return g_path == 1 ? read_persist_prof(out, n) : read_gemm_prof(out, n); }

int rnnt_debug_kernel_times(double* ms, long long* count, int n) {
  if (n < K_NCLASS) return fail(RNNT_ERR_INVALID_ARGUMENT, "need room for %d classes", (int)K_NCLASS);
  CUDA_TRY(cudaDeviceSynchronize());
  for (int c = 0; c < K_NCLASS; ++c) {
    double t = 0.0;
    for (auto& pr : g_pairs[c]) {
      float m = 0.0f;
      cudaEventElapsedTime(&m, pr.first, pr.second);
      t += m;
      g_event_pool.push_back(pr.first);
      g_event_pool.push_back(pr.second);
    }
    ms[c] = t;
    count[c] = static_cast<long long>(g_pairs[c].size());
    g_pairs[c].clear();
  }
  return RNNT_OK;
}

size_t rnnt_fused_workspace_bytes(int B, int Tmax, int Umax, int V, int H) {
  if (check_dims(B, Tmax, Umax, V, H) != RNNT_OK) return 0;
  return make_plan(B, Tmax, Umax, V, H).total;
}

size_t rnnt_fused_state_bytes(int B, int Tmax, int Umax, int V, int H) {
  if (check_dims(B, Tmax, Umax, V, H) != RNNT_OK) return 0;
  return make_plan(B, Tmax, Umax, V, H).state_bytes;
}

size_t rnnt_fused_kept_bytes(int B, int Tmax, int Umax, int V, int H) {
  if (check_dims(B, Tmax, Umax, V, H) != RNNT_OK) return 0;
  return make_plan(B, Tmax, Umax, V, H).kept_bytes;
}

// Whether a pair of calls with a `kept` buffer really keeps the activations: both calls evaluate this the same way.
static bool keeps_activations(const Plan& p, const void* kept, size_t kept_bytes) {
  return kept && p.kept_bytes > 0 && kept_bytes >= p.kept_bytes && g_path == 1 && p.keep_ok && g_keep != 0;
}

static int fused_forward_impl(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                       const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                       int H, int blank, float* loss, void* workspace, size_t workspace_bytes, void* kept,
                       size_t kept_bytes, void* stream) {
  int rc = check_dims(B, Tmax, Umax, V, H);
  if (rc) return rc;
  if (!f || !g || !W || !loss || !workspace || (Umax > 0 && !y)) return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  if (blank < 0 || blank >= V) return fail(RNNT_ERR_INVALID_ARGUMENT, "blank=%d must be in [0, %d]", blank, V - 1);
  rc = check_lens(f_lens_host, y_lens_host, B, Tmax, Umax);
  if (rc) return rc;
  const Plan p = make_plan(B, Tmax, Umax, V, H);
  if (workspace_bytes < p.total)
    return fail(RNNT_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, p.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Ws w{static_cast<uint8_t*>(workspace), p};
  int n_tiles = 0;
  rc = upload_lengths(w, f_lens_host, y_lens_host, &n_tiles, s);
  if (rc) return rc;
  const Lattice L = make_lattice(w, n_tiles);
  JointDims d{V, H, p.Vp, blank, Umax > 0 ? Umax : 1};

  const int nc = chunk_cols(V);
  CUtensorMap tm_h, tm_w;
  rc = make_map(&tm_h, w.at<void>(p.o_h), H, static_cast<uint64_t>(p.slab_tiles) * kTileRows, H, 64, 128);
  if (rc) return rc;
  rc = make_map(&tm_w, W, H, V, H, 64, nc / 2);
  if (rc) return rc;

  if (g_path == 1) {
    const int csize = (nc % 32 == 0 && g_cluster == 4) ? 4 : 2;
    // One accumulator chunk per tile (V <= 256): the MMAs of a tile are shorter than its tanh pass, so the pass is bound by
    // the hgen warps -- run eight of them (two per scheduler) instead of four.
    const int hgen_warps = (g_fwd_hgen_warps == 4 || g_fwd_hgen_warps == 8) ? g_fwd_hgen_warps : ((V + nc - 1) / nc == 1 ? 8 : 4);
    int n_ctas = max_ctas_fwd_persist(csize, hgen_warps);
    if (n_ctas > kMaxPersistCtas) n_ctas = kMaxPersistCtas;
    n_ctas = n_ctas / csize * csize;
    const int n_ptiles = (n_tiles + 1) / 2;
    const int need = (2 * n_ptiles + csize - 1) / csize * csize;
    if (n_ctas > need) n_ctas = need;
    CUtensorMap tm_hs;
    rc = make_map(&tm_hs, w.at<void>(p.o_hs), H, static_cast<uint64_t>(kMaxPersistCtas) * 2 * kTileRows, H, 64, 128);
    if (rc) return rc;
    if (csize == 4) {  // each CTA fetches a quarter of a W chunk
      rc = make_map(&tm_w, W, H, V, H, 64, nc / 4);
      if (rc) return rc;
    }
    CUtensorMap tm_z = tm_hs;   // placeholder when nothing is kept (never dereferenced)
    const bool keep = keeps_activations(p, kept, kept_bytes);
    FwdPArgs pa{};
    if (keep) {
      uint8_t* kb = static_cast<uint8_t*>(kept);
      const uint64_t keep_rows = static_cast<uint64_t>(p.max_tiles) * kTileRows;
      if ((rc = make_map_f16(&tm_z, kb, p.Vp, keep_rows, p.Vp, false))) return rc;
      // the h tiles go straight to their place in the kept buffer and are read back from there as the A operand
      pa.hkeep = reinterpret_cast<__nv_bfloat16*>(kb + p.o_keep_h);
      if ((rc = make_map(&tm_hs, pa.hkeep, H, keep_rows, H, 64, 128))) return rc;
    }
    pa.keep_z = keep ? 1 : 0;
    pa.L = L; pa.dbg = get_gemm_dbg(); pa.csize = csize; pa.hgen_warps = hgen_warps; pa.n_tiles_total = n_tiles; pa.V = V; pa.H = H; pa.nc = nc; pa.n_chunks = (V + nc - 1) / nc;
    pa.k_blocks = (H + 63) / 64; pa.blank = blank; pa.Umax = d.Umax;
    pa.f = static_cast<const __nv_bfloat16*>(f); pa.g = static_cast<const __nv_bfloat16*>(g);
    pa.hscratch = w.at<__nv_bfloat16>(p.o_hs);
    pa.bias = bias; pa.y = y; pa.lse_tile = w.at<float>(p.o_lse); pa.lpb = w.at<float>(p.o_lpb); pa.lpl = w.at<float>(p.o_lpl);
    KLAUNCH(K_FWD, s, launch_fwd_persist(tm_hs, tm_w, tm_z, pa, n_ctas, s));
  }
  FwdArgs a{bias, y, w.at<float>(p.o_lse), w.at<float>(p.o_lpb), w.at<float>(p.o_lpl)};
  for (int t0 = 0; g_path != 1 && t0 < n_tiles; t0 += p.slab_tiles) {
    const int nt = (n_tiles - t0 < p.slab_tiles) ? n_tiles - t0 : p.slab_tiles;
    KLAUNCH(K_HGEN, s, launch_hgen(L, static_cast<const __nv_bfloat16*>(f), static_cast<const __nv_bfloat16*>(g),
                                   w.at<__nv_bfloat16>(p.o_h), t0, nt, H, s));
    KLAUNCH(K_FWD, s, launch_joint_fwd(L, d, tm_h, tm_w, a, t0, nt, nc, s));
  }
  KLAUNCH(K_LATTICE, s, launch_lattice_alpha_beta(L, w.at<float>(p.o_lpb), w.at<float>(p.o_lpl),
                                                  w.at<double>(p.o_alpha), w.at<double>(p.o_beta), loss,
                                                  w.at<float>(p.o_lnpb), w.at<double>(p.o_lnp64), s));
  KLAUNCH(K_COEFS, s, launch_lattice_coefs(L, w.at<float>(p.o_lpb), w.at<float>(p.o_lpl), w.at<double>(p.o_alpha),
                                           w.at<double>(p.o_beta), w.at<double>(p.o_lnp64), w.at<float>(p.o_c1),
                                           w.at<float>(p.o_c2), s));
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

int rnnt_fused_forward(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                       const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                       int H, int blank, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  return fused_forward_impl(f, g, W, bias, y, f_lens_host, y_lens_host, B, Tmax, Umax, V, H, blank, loss, workspace,
                            workspace_bytes, nullptr, 0, stream);
}

int rnnt_fused_forward_keep(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                            const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                            int H, int blank, float* loss, void* workspace, size_t workspace_bytes, void* kept,
                            size_t kept_bytes, void* stream) {
  return fused_forward_impl(f, g, W, bias, y, f_lens_host, y_lens_host, B, Tmax, Umax, V, H, blank, loss, workspace,
                            workspace_bytes, kept, kept_bytes, stream);
}

static int fused_backward_impl(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                        const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                        int H, int blank, const float* grad_loss, float* df, float* dg, float* dW, float* db,
                        void* workspace, size_t workspace_bytes, const void* kept, size_t kept_bytes, void* stream) {
  int rc = check_dims(B, Tmax, Umax, V, H);
  if (rc) return rc;
  if (!f || !g || !W || !grad_loss || !df || !dg || !dW || !db || !workspace || (Umax > 0 && !y))
    return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  if (blank < 0 || blank >= V) return fail(RNNT_ERR_INVALID_ARGUMENT, "blank=%d must be in [0, %d]", blank, V - 1);
  rc = check_lens(f_lens_host, y_lens_host, B, Tmax, Umax);
  if (rc) return rc;
  const Plan p = make_plan(B, Tmax, Umax, V, H);
  if (workspace_bytes < p.total)
    return fail(RNNT_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes, p.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Ws w{static_cast<uint8_t*>(workspace), p};
  const int n_tiles = count_tiles(f_lens_host, y_lens_host, B);
  const Lattice L = make_lattice(w, n_tiles);
  JointDims d{V, H, p.Vp, blank, Umax > 0 ? Umax : 1};

  CUDA_TRY(cudaMemsetAsync(df, 0, sizeof(float) * static_cast<size_t>(B) * Tmax * H, s));
  CUDA_TRY(cudaMemsetAsync(dg, 0, sizeof(float) * static_cast<size_t>(B) * (Umax + 1) * H, s));
  CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * static_cast<size_t>(V) * H, s));
  CUDA_TRY(cudaMemsetAsync(db, 0, sizeof(float) * V, s));

  const int nc_v = chunk_cols(V), nc_h = chunk_cols(H);
  const bool keep = keeps_activations(p, kept, kept_bytes);
  const int n_cons = keep ? p.Ck : p.C, n_kg = keep ? p.KGk : p.KG;
  const bool mega = g_path == 1 && p.mega_ok && max_ctas_bwd_mega(2) >= 2 * (p.P + p.C);
  const uint64_t wt_rows = align_up(H, 32);
  // the mega-kernel wants the rows of W^T permuted inside groups of 32 (fragment layout of its dh epilogue)
  KLAUNCH(K_MISC, s, launch_transpose_w(static_cast<const __nv_bfloat16*>(W), w.at<__nv_bfloat16>(p.o_wt), V, H, p.Vp, mega ? 1 : 0, s));
  if (mega) {
    // 4-clusters (operand multicast between two pairs of one role): both roles must start on a cluster boundary,
    // both W chunk widths must split into quarters, and only the co-resident capacity for 4-clusters (132 of 148
    // CTAs on this part) can be used, so the producers give up the difference.
    int P = keep ? p.Pk : p.P;
    int csize = 2;
    if (!keep && g_cluster_bwd == 4 && p.C % 2 == 0 && nc_v % 32 == 0 && nc_h % 32 == 0) {
      int P4 = max_ctas_bwd_mega(4) / 2 - p.C;
      if (P4 > p.P) P4 = p.P;
      P4 &= ~1;
      if (P4 >= 2) { P = P4; csize = 4; }
    }
    const uint64_t ring_rows = static_cast<uint64_t>(P) * p.NS * 2 * kTileRows;
    const int wdiv = csize == 4 ? 4 : 2;
    CUtensorMap tm_h, tm_w, tm_dz, tm_wt, tm_dz_mn, tm_h_mn, tm_dz_st;
    if ((rc = make_map(&tm_dz_st, w.at<void>(p.o_dzring), p.Vp, ring_rows, p.Vp, 64, 32))) return rc;
    if ((rc = make_map(&tm_h, w.at<void>(p.o_hring), H, ring_rows, H, 64, 128))) return rc;
    if ((rc = make_map(&tm_w, W, H, V, H, 64, nc_v / wdiv))) return rc;
    if ((rc = make_map(&tm_dz, w.at<void>(p.o_dzring), p.Vp, ring_rows, p.Vp, 64, 128))) return rc;
    if ((rc = make_map(&tm_wt, w.at<void>(p.o_wt), p.Vp, wt_rows, p.Vp, 64, nc_h / wdiv))) return rc;
    if ((rc = make_map(&tm_dz_mn, w.at<void>(p.o_dzring), p.Vp, ring_rows, p.Vp, 64, 64))) return rc;
    if ((rc = make_map(&tm_h_mn, w.at<void>(p.o_hring), H, ring_rows, H, 64, 64))) return rc;
    const size_t n_flags = static_cast<size_t>(kMaxPersistCtas / 2) * kMaxRingSlots;
    CUDA_TRY(cudaMemsetAsync(w.at<unsigned>(p.o_flags), 0, sizeof(unsigned) * 2 * n_flags, s));
    BwdPArgs a{};
    a.L = L; a.dbg = get_gemm_dbg(); a.csize = csize; a.cons_share = (csize == 4 && p.n_vt % 2 == 0) ? 1 : 0;
    a.n_tiles_total = n_tiles; a.V = V; a.H = H; a.Vp = p.Vp;
    if (g_prune >= 0) {   // walk only the tiles whose cells carry occupancy (exact for eps = 0, the default)
      const float eps = g_prune_log2_eps <= -1000 ? 0.0f : exp2f(static_cast<float>(g_prune_log2_eps));
      KLAUNCH(K_MISC, s, launch_tile_activity(L, w.at<float>(p.o_c1), w.at<float>(p.o_c2), grad_loss, eps, w.at<int>(p.o_tflags), s));
      KLAUNCH(K_MISC, s, launch_compact_tiles(w.at<int>(p.o_tflags), n_tiles, w.at<int>(p.o_active), w.at<int>(p.o_nactive), s));
      a.active_tiles = w.at<int>(p.o_active); a.n_active = w.at<int>(p.o_nactive);
    }
    a.nc_v = nc_v; a.n_chunks_v = (V + nc_v - 1) / nc_v; a.kb_h = (H + 63) / 64;
    a.nc_h = nc_h; a.n_chunks_h = (H + nc_h - 1) / nc_h; a.kb_v = p.Vp / 64;
    a.blank = blank; a.Umax = d.Umax;
    a.P = P; a.C = n_cons; a.KG = n_kg; a.NS = p.NS; a.n_vt = p.n_vt; a.n_ht = p.n_ht; a.n_out = p.n_out;
    a.f = static_cast<const __nv_bfloat16*>(f); a.g = static_cast<const __nv_bfloat16*>(g);
    a.h_ring = w.at<__nv_bfloat16>(p.o_hring); a.dz_ring = w.at<__nv_bfloat16>(p.o_dzring);
    CUtensorMap tm_zl = tm_dz;   // placeholder when the logits are recomputed (never dereferenced)
    if (keep) {
      const uint8_t* kb = static_cast<const uint8_t*>(kept);
      const uint64_t keep_rows = static_cast<uint64_t>(p.max_tiles) * kTileRows;
      a.zlog = reinterpret_cast<const uint16_t*>(kb);
      a.hkeep = reinterpret_cast<const __nv_bfloat16*>(kb + p.o_keep_h);
      a.zcols = a.n_chunks_v * nc_v < p.Vp ? a.n_chunks_v * nc_v : p.Vp;
      if ((rc = make_map_f16(&tm_zl, kb, p.Vp, keep_rows, p.Vp, true))) return rc;
      if ((rc = make_map(&tm_h_mn, a.hkeep, H, keep_rows, H, 64, 64))) return rc;   // the consumers read h where the forward pass left it
    }
    a.bias = bias; a.y = y; a.lse_tile = w.at<float>(p.o_lse); a.lpb = w.at<float>(p.o_lpb); a.lpl = w.at<float>(p.o_lpl);
    a.c1 = w.at<float>(p.o_c1); a.c2 = w.at<float>(p.o_c2); a.grad_loss = grad_loss;
    a.db = db; a.df = df; a.dg = dg; a.dW = dW;
    a.ready = w.at<unsigned>(p.o_flags); a.done = w.at<unsigned>(p.o_flags) + n_flags;
    KLAUNCH(K_BWD_MEGA, s, launch_bwd_mega(tm_h, tm_w, tm_dz, tm_wt, tm_dz_mn, tm_h_mn, tm_dz_st, tm_zl, a, 2 * (P + n_cons), s));
    CUDA_TRY(cudaGetLastError());
    return RNNT_OK;
  }
  const uint64_t slab_rows = static_cast<uint64_t>(p.slab_tiles) * kTileRows;
  CUtensorMap tm_h, tm_w, tm_dz, tm_wt, tm_dz_mn, tm_h_mn;
  if ((rc = make_map(&tm_h, w.at<void>(p.o_h), H, slab_rows, H, 64, 128))) return rc;
  if ((rc = make_map(&tm_w, W, H, V, H, 64, nc_v / 2))) return rc;
  if ((rc = make_map(&tm_dz, w.at<void>(p.o_dz), p.Vp, slab_rows, p.Vp, 64, 128))) return rc;
  if ((rc = make_map(&tm_wt, w.at<void>(p.o_wt), p.Vp, wt_rows, p.Vp, 64, nc_h / 2))) return rc;
  if ((rc = make_map(&tm_dz_mn, w.at<void>(p.o_dz), p.Vp, slab_rows, p.Vp, 64, 64))) return rc;
  if ((rc = make_map(&tm_h_mn, w.at<void>(p.o_h), H, slab_rows, H, 64, 64))) return rc;

  DzArgs za{bias, y, w.at<float>(p.o_lse), w.at<float>(p.o_lpb), w.at<float>(p.o_lpl), w.at<float>(p.o_c1),
            w.at<float>(p.o_c2), grad_loss, db};
  DhArgs ha{w.at<__nv_bfloat16>(p.o_h), df, dg};
  const int n_sm = sm_count();
  for (int t0 = 0; t0 < n_tiles; t0 += p.slab_tiles) {
    const int nt = (n_tiles - t0 < p.slab_tiles) ? n_tiles - t0 : p.slab_tiles;
    KLAUNCH(K_HGEN, s, launch_hgen(L, static_cast<const __nv_bfloat16*>(f), static_cast<const __nv_bfloat16*>(g),
                                   w.at<__nv_bfloat16>(p.o_h), t0, nt, H, s));
    KLAUNCH(K_DZ, s, launch_joint_dz(L, d, tm_h, tm_w, tm_dz, za, t0, nt, nc_v, s));
    KLAUNCH(K_DH, s, launch_joint_dh(L, d, tm_dz, tm_wt, ha, t0, nt, nc_h, s));
    KLAUNCH(K_DW, s, launch_joint_dw(d, tm_dz_mn, tm_h_mn, dW, nt, n_sm, s));
  }
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

int rnnt_fused_backward(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                        const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                        int H, int blank, const float* grad_loss, float* df, float* dg, float* dW, float* db,
                        void* workspace, size_t workspace_bytes, void* stream) {
  return fused_backward_impl(f, g, W, bias, y, f_lens_host, y_lens_host, B, Tmax, Umax, V, H, blank, grad_loss, df, dg, dW,
                             db, workspace, workspace_bytes, nullptr, 0, stream);
}

int rnnt_fused_backward_kept(const void* f, const void* g, const void* W, const float* bias, const int32_t* y,
                             const int32_t* f_lens_host, const int32_t* y_lens_host, int B, int Tmax, int Umax, int V,
                             int H, int blank, const float* grad_loss, float* df, float* dg, float* dW, float* db,
                             void* workspace, size_t workspace_bytes, const void* kept, size_t kept_bytes,
                             void* stream) {
  return fused_backward_impl(f, g, W, bias, y, f_lens_host, y_lens_host, B, Tmax, Umax, V, H, blank, grad_loss, df, dg, dW,
                             db, workspace, workspace_bytes, kept, kept_bytes, stream);
}

size_t rnnt_lattice_workspace_bytes(int B, int Tmax, int Umax) {
  if (B < 1 || Tmax < 1 || Umax < 0) return 0;
  const size_t cells = static_cast<size_t>(B) * (Tmax + Umax + 1) * (Umax + 1);
  return 3 * 1024 + align_up(sizeof(int) * (B + 1), 1024) * 3 + 4 * align_up(sizeof(float) * cells, 1024) +
         2 * align_up(sizeof(double) * cells, 1024) + align_up(sizeof(float) * B, 1024) + align_up(sizeof(double) * B, 1024);
}

int rnnt_lattice_forward(const float* lp_blank, const float* lp_label, const int32_t* f_lens_host,
                         const int32_t* y_lens_host, int B, int Tmax, int Umax, float* loss, float* c_blank,
                         float* c_label, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || Tmax < 1 || Umax < 0) return fail(RNNT_ERR_INVALID_ARGUMENT, "B=%d Tmax=%d Umax=%d out of range", B, Tmax, Umax);
  if (Umax + 1 > lattice_max_columns())
    return fail(RNNT_ERR_UNSUPPORTED, "Umax+1=%d must be <= %d", Umax + 1, lattice_max_columns());
  if (!lp_blank || !lp_label || !loss || !c_blank || !c_label || !workspace)
    return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  int rc = check_lens(f_lens_host, y_lens_host, B, Tmax, Umax);
  if (rc) return rc;
  if (workspace_bytes < rnnt_lattice_workspace_bytes(B, Tmax, Umax))
    return fail(RNNT_ERR_WORKSPACE_TOO_SMALL, "workspace %zu bytes < required %zu", workspace_bytes,
                rnnt_lattice_workspace_bytes(B, Tmax, Umax));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const int U1 = Umax + 1, D = Tmax + U1;
  const size_t cells = static_cast<size_t>(B) * D * U1;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return base + r; };
  int* d_fl = reinterpret_cast<int*>(take(sizeof(int) * B));
  int* d_yl = reinterpret_cast<int*>(take(sizeof(int) * B));
  float* lpb = reinterpret_cast<float*>(take(sizeof(float) * cells));
  float* lpl = reinterpret_cast<float*>(take(sizeof(float) * cells));
  double* al = reinterpret_cast<double*>(take(sizeof(double) * cells));
  double* be = reinterpret_cast<double*>(take(sizeof(double) * cells));
  float* c1 = reinterpret_cast<float*>(take(sizeof(float) * cells));
  float* c2 = reinterpret_cast<float*>(take(sizeof(float) * cells));
  float* lnpb = reinterpret_cast<float*>(take(sizeof(float) * B));
  double* lnp64 = reinterpret_cast<double*>(take(sizeof(double) * B));
  CUDA_TRY(cudaMemcpyAsync(d_fl, f_lens_host, sizeof(int) * B, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d_yl, y_lens_host, sizeof(int) * B, cudaMemcpyHostToDevice, s));
  Lattice L{};
  L.tile_prefix = nullptr; L.f_lens = d_fl; L.y_lens = d_yl; L.B = B; L.Tmax = Tmax; L.U1max = U1; L.D = D;
  KLAUNCH(K_MISC, s, launch_nat_to_diag(L, lp_blank, lp_label, lpb, lpl, s));
  KLAUNCH(K_LATTICE, s, launch_lattice_alpha_beta(L, lpb, lpl, al, be, loss, lnpb, lnp64, s));
  KLAUNCH(K_COEFS, s, launch_lattice_coefs(L, lpb, lpl, al, be, lnp64, c1, c2, s));
  KLAUNCH(K_MISC, s, launch_diag_to_nat(L, c1, c2, c_blank, c_label, s));
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

int rnnt_greedy_joint_argmax(const void* f, const void* g, const void* W, const float* bias, const int32_t* t_idx,
                             int32_t* out_k, int B, int Tmax, int V, int H, void* stream) {
  if (B < 1 || Tmax < 1 || V < 1) return fail(RNNT_ERR_INVALID_ARGUMENT, "B=%d Tmax=%d V=%d out of range", B, Tmax, V);
  if (H < 8 || H % 8 != 0) return fail(RNNT_ERR_UNSUPPORTED, "H=%d must be a positive multiple of 8", H);
  if (H > 8192) return fail(RNNT_ERR_UNSUPPORTED, "H=%d must be <= 8192", H);
  if (!f || !g || !W || !t_idx || !out_k) return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  KLAUNCH(K_MISC, static_cast<cudaStream_t>(stream),
          launch_greedy_argmax(static_cast<const __nv_bfloat16*>(f), static_cast<const __nv_bfloat16*>(g),
                               static_cast<const __nv_bfloat16*>(W), bias, t_idx, out_k, B, Tmax, V, H,
                               static_cast<cudaStream_t>(stream)));
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

int rnnt_greedy_step(const void* f, const float* g, const void* W, const float* bias, const int32_t* lens,
                     int32_t* t_cur, int32_t* emitted, int32_t* n_sym, int32_t* sym, int sym_cap, int32_t* is_sym,
                     int32_t* label, int32_t* active, int B, int Tmax, int V, int H, int blank, int max_symbols,
                     void* stream) {
  if (B < 1 || Tmax < 1 || V < 1 || sym_cap < 1 || max_symbols < 1)
    return fail(RNNT_ERR_INVALID_ARGUMENT, "B=%d Tmax=%d V=%d sym_cap=%d max_symbols=%d out of range", B, Tmax, V, sym_cap, max_symbols);
  if (H < 8 || H % 8 != 0 || H > 8192) return fail(RNNT_ERR_UNSUPPORTED, "H=%d must be a multiple of 8 in [8, 8192]", H);
  if (blank < 0 || blank >= V) return fail(RNNT_ERR_INVALID_ARGUMENT, "blank=%d must be in [0, %d]", blank, V - 1);
  if (!f || !g || !W || !lens || !t_cur || !emitted || !n_sym || !sym || !is_sym || !label || !active)
    return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  KLAUNCH(K_MISC, static_cast<cudaStream_t>(stream),
          launch_greedy_step(static_cast<const __nv_bfloat16*>(f), g, static_cast<const __nv_bfloat16*>(W), bias, lens, t_cur,
                             emitted, n_sym, sym, sym_cap, is_sym, label, active, B, Tmax, V, H, blank, max_symbols,
                             static_cast<cudaStream_t>(stream)));
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

size_t rnnt_greedy_decode_stack_workspace_bytes(int B, int V, int H, int Hp, int n_layers) {
  const CDecPlan c = make_cdec_plan(B, V, H, Hp, g_decode_cluster, n_layers);
  size_t a = 0;
  if (n_layers == 1) { const DecPlan d = make_dec_plan(B, V, H, Hp); a = d.ok ? d.w_total : 0; }
  const size_t b = c.ok ? c.w_total : 0;
  return a > b ? a : b;
}

size_t rnnt_greedy_decode_workspace_bytes(int B, int V, int H, int Hp) {
  return rnnt_greedy_decode_stack_workspace_bytes(B, V, H, Hp, 1);
}

int rnnt_greedy_decode_lstm(const void* f, const int32_t* lens, const void* W, const float* bias, const float* gate_table,
                            const void* W_hh, const void* W_proj, const float* bias_proj, int B, int Tmax, int V, int H,
                            int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap, int32_t* n_sym, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return rnnt_greedy_decode_lstm_stack(f, lens, W, bias, gate_table, W_hh, 1, nullptr, nullptr, W_proj, bias_proj, B, Tmax, V,
                                       H, Hp, blank, max_symbols, sym, sym_cap, n_sym, workspace, workspace_bytes, stream);
}

static int decode_stack_impl(int cell, const void* f, const int32_t* lens, const void* W, const float* bias,
                             const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                             const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax, int V,
                             int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap, int32_t* n_sym,
                             void* workspace, size_t workspace_bytes, void* stream);

int rnnt_greedy_decode_lstm_stack(const void* f, const int32_t* lens, const void* W, const float* bias,
                                  const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                                  const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax, int V,
                                  int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap, int32_t* n_sym,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  return decode_stack_impl(0, f, lens, W, bias, gate_table, W_hh, n_layers, W_upper, bias_upper, W_proj, bias_proj, B, Tmax, V,
                           H, Hp, blank, max_symbols, sym, sym_cap, n_sym, workspace, workspace_bytes, stream);
}

int rnnt_greedy_decode_gru_stack(const void* f, const int32_t* lens, const void* W, const float* bias,
                                 const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                                 const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax, int V,
                                 int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap, int32_t* n_sym,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  return decode_stack_impl(1, f, lens, W, bias, gate_table, W_hh, n_layers, W_upper, bias_upper, W_proj, bias_proj, B, Tmax, V,
                           H, Hp, blank, max_symbols, sym, sym_cap, n_sym, workspace, workspace_bytes, stream);
}

static int decode_stack_impl(int cell, const void* f, const int32_t* lens, const void* W, const float* bias,
                             const float* gate_table, const void* W_hh, int n_layers, const void* W_upper,
                             const float* bias_upper, const void* W_proj, const float* bias_proj, int B, int Tmax, int V,
                             int H, int Hp, int blank, int max_symbols, int32_t* sym, int sym_cap, int32_t* n_sym,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || Tmax < 1 || V < 1 || sym_cap < 1 || max_symbols < 1 || n_layers < 1)
    return fail(RNNT_ERR_INVALID_ARGUMENT, "B=%d Tmax=%d V=%d sym_cap=%d max_symbols=%d n_layers=%d out of range", B, Tmax, V,
                sym_cap, max_symbols, n_layers);
  if (blank < 0 || blank >= V) return fail(RNNT_ERR_INVALID_ARGUMENT, "blank=%d must be in [0, %d]", blank, V - 1);
  if (!f || !lens || !W || !gate_table || !W_hh || !W_proj || !sym || !n_sym || !workspace ||
      (n_layers > 1 && (!W_upper || !bias_upper)))
    return fail(RNNT_ERR_INVALID_ARGUMENT, "NULL pointer argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const CDecPlan c = make_cdec_plan(B, V, H, Hp, g_decode_cluster, n_layers);
  const bool want_cluster = g_decode_variant == 1 || n_layers > 1 || cell != 0;   // the grid-barrier schedule: one LSTM layer
  if (want_cluster && c.ok && workspace_bytes >= c.w_total && max_clusters_greedy_decode(c.smem, c.C) >= 1) {
    __nv_bfloat16* whh_perm = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wup_perm = reinterpret_cast<__nv_bfloat16*>(ws + c.w_wup);
    const int rows_l = c.C * 4 * c.up;
    KLAUNCH(K_MISC, s, launch_permute_whh_cluster(static_cast<const __nv_bfloat16*>(W_hh), whh_perm, Hp, c.up, rows_l, s));
    if (n_layers > 1)
      KLAUNCH(K_MISC, s, launch_permute_wup_cluster(static_cast<const __nv_bfloat16*>(W_upper), wup_perm, Hp, c.up, c.C,
                                                    c.kbHp, (n_layers - 1) * rows_l, s));
    CUtensorMap tm_wj, tm_wl, tm_wu, tm_wp;
    int rc;
    if ((rc = make_map(&tm_wj, W, H, V, H, 64, 128))) return rc;
    if ((rc = make_map(&tm_wl, whh_perm, Hp, rows_l, Hp, 64, 128))) return rc;
    if (n_layers > 1) {
      if ((rc = make_map(&tm_wu, wup_perm, 2 * c.kbHp * 64, static_cast<uint64_t>(n_layers - 1) * rows_l, 2 * c.kbHp * 64, 64, 128)))
        return rc;
    } else {
      tm_wu = tm_wl;   // never dereferenced
    }
    if ((rc = make_map(&tm_wp, W_proj, Hp, H, Hp, 64, 128))) return rc;
    ClusterDecodeArgs a{};
    a.B = B; a.Tmax = Tmax; a.V = V; a.H = H; a.Hp = Hp; a.blank = blank; a.S = max_symbols; a.sym_cap = sym_cap;
    a.max_steps = Tmax * max_symbols + 1;
    a.C = c.C; a.RJ = c.RJ; a.RP = c.RP; a.up = c.up; a.mtJ = c.mtJ; a.mtP = c.mtP; a.mtL = c.mtL; a.kbH = c.kbH; a.kbHp = c.kbHp;
    a.n_stages = c.n_stages; a.tmem_cols = c.tmem_cols; a.NL = n_layers; a.cell = cell;
    a.o_hj = c.o_hj; a.o_layers = c.o_layers; a.layer_stride = c.layer_stride; a.o_gates = c.o_gates;
    a.o_amax = c.o_amax; a.o_part = c.o_part; a.o_state = c.o_state; a.o_bars = c.o_bars;
    a.f = static_cast<const __nv_bfloat16*>(f); a.lens = lens; a.bias_j = bias; a.table = gate_table; a.bias_up = bias_upper;
    a.bias_p = bias_proj;
    a.sym = sym; a.n_sym = n_sym; a.prof = g_decode_prof; a.l_late = g_decode_l_late;
    {
      // Spare TMEM columns hold weights for the whole decode (32 columns per 64-wide k-block of a 128-row tile; A operand
      // of the TS-form MMA): first the projection tile, then leading k-blocks of the first vocabulary tile.  (Measured
      // the other way round -- 8 k-blocks of W, 4 of W_p, so that the streamed rest of W fits the ring: 23.7 -> 25.7 ms
      // at warm clocks; the projection sits right behind the h exchange and gains more from needing no weights at all.)
      const int acc_cols = 32 * (n_layers * c.mtL + c.mtP + c.mtJ);
      int budget = g_decode_res ? (512 - acc_cols) / 32 : 0;
      if (budget < 0) budget = 0;
      a.res_p = c.mtP == 1 ? (budget < c.kbHp ? budget : c.kbHp) : 0;
      a.res_j = budget - a.res_p < c.kbH ? budget - a.res_p : c.kbH;
      a.res_col = acc_cols;
      a.res_j_col = acc_cols + 32 * a.res_p;
      a.w_proj = static_cast<const __nv_bfloat16*>(W_proj);
      a.w_joint = static_cast<const __nv_bfloat16*>(W);
      if (a.res_p || a.res_j) a.tmem_cols = 512;
    }
    cudaError_t e = cudaSuccess;
    KLAUNCH(K_MISC, s, e = launch_greedy_decode_cluster(tm_wj, tm_wl, tm_wu, tm_wp, a, (B + 15) / 16, c.smem, s));
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RNNT_ERR_CUDA, "greedy decode (cluster) launch -> %s", cudaGetErrorString(e)); }
    CUDA_TRY(cudaGetLastError());
    return RNNT_OK;
  }
  if (n_layers > 1 || cell != 0)
    return fail(RNNT_ERR_UNSUPPORTED, "fused greedy decode does not cover B=%d V=%d H=%d Hp=%d with %d %s layers", B, V, H, Hp,
                n_layers, cell ? "GRU" : "LSTM");
  const DecPlan d = make_dec_plan(B, V, H, Hp);
  if (!d.ok) return fail(RNNT_ERR_UNSUPPORTED, "fused greedy decode does not cover B=%d V=%d H=%d Hp=%d", B, V, H, Hp);
  if (workspace_bytes < d.w_total)
    return fail(RNNT_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu bytes", workspace_bytes, d.w_total);
  const int G = max_ctas_greedy_decode(d.smem);
  if (G < d.nslJ || G < d.nslL || G < d.nslP)
    return fail(RNNT_ERR_UNSUPPORTED, "fused greedy decode needs %d co-resident CTAs, the device offers %d",
                std::max(d.nslJ, std::max(d.nslL, d.nslP)), G);
  CUDA_TRY(cudaMemsetAsync(ws, 0, d.w_whh, s));  // barrier counter, argmax keys, hj, both h buffers (h_0 = 0)
  __nv_bfloat16* whh_perm = reinterpret_cast<__nv_bfloat16*>(ws + d.w_whh);
  KLAUNCH(K_MISC, s, launch_permute_whh(static_cast<const __nv_bfloat16*>(W_hh), whh_perm, Hp, d.nu, d.nslL * d.nL, s));
  CUtensorMap tm_hj, tm_hbuf, tm_wj, tm_wl, tm_wp;
  int rc;
  if ((rc = make_map(&tm_hj, ws + d.w_hj, H, d.Bp, H, 64, 128))) return rc;
  if ((rc = make_map(&tm_hbuf, ws + d.w_hbuf, Hp, 2 * static_cast<uint64_t>(d.Bp), Hp, 64, 128))) return rc;
  if ((rc = make_map(&tm_wj, W, H, V, H, 64, d.nJ))) return rc;
  if ((rc = make_map(&tm_wl, whh_perm, Hp, static_cast<uint64_t>(d.nslL) * d.nL, Hp, 64, d.nL))) return rc;
  if ((rc = make_map(&tm_wp, W_proj, Hp, H, Hp, 64, d.nP))) return rc;
  DecodeArgs a{};
  a.B = B; a.Bp = d.Bp; a.Tmax = Tmax; a.V = V; a.H = H; a.Hp = Hp; a.blank = blank; a.S = max_symbols; a.sym_cap = sym_cap;
  a.max_steps = Tmax * max_symbols + 1;
  a.nJ = d.nJ; a.nslJ = d.nslJ; a.kbJ = d.kbJ; a.nu = d.nu; a.nL = d.nL; a.nslL = d.nslL; a.kbL = d.kbL;
  a.nP = d.nP; a.nslP = d.nslP; a.kbP = d.kbP; a.n_stages = d.n_stages;
  a.o_wj = d.o_wj; a.o_wl = d.o_wl; a.o_wp = d.o_wp; a.o_c = d.o_c; a.o_h = d.o_h; a.o_g = d.o_g; a.o_state = d.o_state;
  a.o_bars = d.o_bars;
  a.f = static_cast<const __nv_bfloat16*>(f); a.lens = lens; a.bias_j = bias; a.table = gate_table; a.bias_p = bias_proj;
  a.hj = reinterpret_cast<__nv_bfloat16*>(ws + d.w_hj);
  a.hbuf = reinterpret_cast<__nv_bfloat16*>(ws + d.w_hbuf);
  a.amax = reinterpret_cast<unsigned long long*>(ws + d.w_amax);
  a.gbar = reinterpret_cast<unsigned*>(ws + d.w_gbar);
  a.sym = sym; a.n_sym = n_sym;
  cudaError_t e = cudaSuccess;
  KLAUNCH(K_MISC, s, e = launch_greedy_decode(tm_hj, tm_hbuf, tm_wj, tm_wl, tm_wp, a, G, d.smem, s));
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RNNT_ERR_CUDA, "greedy decode launch -> %s", cudaGetErrorString(e)); }
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

int rnnt_debug_copy_stats(const void* workspace, int B, int Tmax, int Umax, int V, int H, float* lp_blank,
                          float* lp_label, float* c_blank, float* c_label, float* lnp_beta, void* stream) {
  int rc = check_dims(B, Tmax, Umax, V, H);
  if (rc) return rc;
  const Plan p = make_plan(B, Tmax, Umax, V, H);
  Ws w{static_cast<uint8_t*>(const_cast<void*>(workspace)), p};
  const Lattice L = make_lattice(w, 0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (lp_blank && lp_label) launch_diag_to_nat(L, w.at<float>(p.o_lpb), w.at<float>(p.o_lpl), lp_blank, lp_label, s);
  if (c_blank && c_label) launch_diag_to_nat(L, w.at<float>(p.o_c1), w.at<float>(p.o_c2), c_blank, c_label, s);
  if (lnp_beta) CUDA_TRY(cudaMemcpyAsync(lnp_beta, w.at<float>(p.o_lnpb), sizeof(float) * B, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaGetLastError());
  return RNNT_OK;
}

}  // extern "C"
