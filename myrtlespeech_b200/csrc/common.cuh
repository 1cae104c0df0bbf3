// Shared device-side structures for the RNN-T transducer-head kernels.
//
// Row order.  The joint is evaluated on "lattice rows" (b, t, u).  Rows are grouped into tiles of
// 128 = kTT x kTU lattice cells (16 frames x 8 label positions) so that the backward reductions
// df = sum_u and dg = sum_t collapse 8x / 16x on chip before they touch HBM.  Tiles never straddle
// utterances; `tile_prefix[b]` is the first tile of utterance b.
//
// Diagonal layout.  Per-cell lattice scalars (lp_blank, lp_label, alpha, beta, c1, c2) are stored
// as X[b][d = t+u][u] so the anti-diagonal wavefront of the alpha/beta recurrences reads and writes
// contiguous memory.
#pragma once
#include <stdint.h>

namespace rnnt {

constexpr int kTT = 16;           // frames per tile
constexpr int kTU = 8;            // label positions per tile
constexpr int kTileRows = 128;    // kTT * kTU
constexpr float kNeg = -1.0e30f;  // log(0) stand-in that survives additions

struct Lattice {
  const int* tile_prefix;  // [B+1] device
  const int* f_lens;       // [B]   device
  const int* y_lens;       // [B]   device
  int B;
  int Tmax;
  int U1max;  // Umax + 1
  int D;      // diagonals allocated per utterance (Tmax + U1max)
  int n_tiles_total;  // tiles in the batch (== tile_prefix[B])
};

struct TileInfo {
  int b, t0, u0, T, U;  // U = y_len (lattice has U+1 columns)
};

__device__ __forceinline__ TileInfo decode_tile(const Lattice& L, int gtile) {
  int lo = 0, hi = L.B;  // invariant: prefix[lo] <= gtile < prefix[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (L.tile_prefix[mid] <= gtile) lo = mid; else hi = mid;
  }
  TileInfo ti;
  ti.b = lo;
  ti.T = L.f_lens[lo];
  ti.U = L.y_lens[lo];
  const int n_ub = (ti.U + 1 + kTU - 1) / kTU;
  const int local = gtile - L.tile_prefix[lo];
  ti.t0 = (local / n_ub) * kTT;
  ti.u0 = (local % n_ub) * kTU;
  return ti;
}

// The persistent kernels walk tiles in increasing order, so the utterance of the next tile is the current one or a later
// one: a cursor that remembers (b, prefix[b], prefix[b+1], T_b, U_b) replaces the binary search -- five dependent loads
// that miss L1 on every tile, because the gpu-scope fences of the producer warps invalidate it -- by loads only where an
// utterance boundary is crossed (measured: 3.5 k of the 12 k cycles the V = 29 forward pass spent per tile).
struct TileCursor {
  int b, lo, hi, T, U;
  __device__ __forceinline__ void init(const Lattice& L) {
    b = 0; lo = 0; hi = L.tile_prefix[1]; T = L.f_lens[0]; U = L.y_lens[0];
  }
  __device__ __forceinline__ TileInfo at(const Lattice& L, int gtile) {
    if (gtile >= hi && b + 1 < L.B) {
      do { ++b; lo = hi; hi = L.tile_prefix[b + 1]; } while (gtile >= hi && b + 1 < L.B);
      T = L.f_lens[b]; U = L.y_lens[b];
    }
    TileInfo ti;
    ti.b = b; ti.T = T; ti.U = U;
    const int n_ub = (U + 1 + kTU - 1) / kTU;
    const int local = gtile - lo;
    ti.t0 = (local / n_ub) * kTT;
    ti.u0 = (local % n_ub) * kTU;
    return ti;
  }
};

__device__ __forceinline__ size_t diag_index(const Lattice& L, int b, int t, int u) {
  return (static_cast<size_t>(b) * L.D + (t + u)) * L.U1max + u;
}

}  // namespace rnnt
