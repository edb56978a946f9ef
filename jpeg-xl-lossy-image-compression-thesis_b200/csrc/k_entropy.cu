// K8-K10 — AC coefficient tokenisation + per-context histograms, histogram clustering,
// per-cluster ANS tables, and the per-group reverse rANS encoder (stages U6-U8; libjxl
// enc_entropy_coder.cc TokenizeCoefficients, ac_context.h, enc_cluster.cc, enc_ans.cc
// [UPSTREAM]; algorithmic choices in DESIGN.md "Entropy stage").
//
//  k_tokenize    four CTAs per 256x256 AC group: token counts per block-channel from (nzeros, last
//                non-zero scan position), CTA-wide exclusive scan (recomputed by each of the four), then
//                rounds of 32 block-channel entries drawn from a CTA counter: every lane prefetches one
//                entry's metadata and first 16 scan positions, the warp walks four entries side by side
//                (8 lanes each) with ballot/popc for the running non-zero count; tokens leave as 32-bit
//                stores, histogram bins through shared-memory counters (symbols 0/1) or global REDs.
//  k_cluster     ONE thread-block cluster of 8 CTAs: farthest-point seeding on an integer
//                entropy distance; the per-round argmax travels through distributed shared
//                memory (one cluster barrier per round instead of a grid-wide sync).
//  k_ans_tables  one warp per cluster histogram: normalise to 4096, code the histogram header,
//                build the alias table and the encoder's reverse map.
//  k_ans_groups  one warp per AC group, software-pipelined over chunks of 32 tokens: while the serial state
//                chain of chunk c runs on operands broadcast from shared memory, the same warp looks up the
//                cluster / symbol info of chunk c + 1 and places the bits of chunk c - 1 (prefix sum, OR into
//                a staging window, coalesced stores); the stream grows backwards in the group's arena.
#include "entropy.cuh"
#include "kernels.h"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace jxlb {

__constant__ uint8_t c_covered_x[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
__constant__ uint8_t c_covered_y[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
__constant__ uint8_t c_strategy_order[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
__constant__ uint8_t c_block_ctx_map[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
                                            7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};
__constant__ uint8_t c_freq_ctx[64] = {0,  0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17,
                                       18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25,
                                       26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
__constant__ uint8_t c_nnz_ctx[64] = {0,   0,   31,  62,  62,  93,  93,  93,  93,  123, 123, 123, 123, 152, 152, 152,
                                      152, 152, 152, 152, 152, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180,
                                      180, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
                                      206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

// ------------------------------------------------------------------------------------------ K8
constexpr int kTokSplit = 4;   // CTAs per AC group (each recomputes the group's scan, emits 1/4 of the entries)

__global__ void __launch_bounds__(256) k_tokenize(const uint8_t* __restrict__ acs, const uint8_t* __restrict__ nzeros,
                                                  const uint16_t* __restrict__ nzcount, const uint16_t* __restrict__ lastk,
                                                  const int16_t* __restrict__ coeffs, FrameDim fd,
                                                  uint32_t* __restrict__ tokens, uint32_t* __restrict__ token_counts,
                                                  uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_off[3072 + 1];
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_next;
  // first 16 scan positions of each of the warp's 32 entries, fetched by the entry's own lane together with its
  // metadata (two 16-byte loads in flight per lane) so that the walk below reads them from shared memory instead of
  // waiting for one dependent 2-byte global load per round
  __shared__ __align__(16) int16_t s_coef[8][32][16];
  // CTA-private counters of symbols 0 and 1 of every context: those bins take most of the increments
  // (zero coefficients, +-1) and a few of them are so hot that global atomics on them serialise in L2
  // (one word per context: symbol 0 in the low half, symbol 1 in the high half — a CTA emits at most 256 blocks x 3 x 64
  // = 49 152 tokens, so neither half can overflow; 30 KB instead of 59 KB lets a fourth CTA live on the SM)
  extern __shared__ uint32_t s_hot[];   // [kNumAcContexts]
  const int g = blockIdx.x, part = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < kNumAcContexts; i += 256) s_hot[i] = 0;
  if (t == 0) s_next = 0;
  const int gx0 = (g % fd.gxs) * 32, gy0 = (g / fd.gxs) * 32;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  // ---- token count per (block, slot): 12 consecutive entries per thread = 4 blocks
  uint32_t cnt[12];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int blk = t * 4 + i, lx = blk & 31, ly = blk >> 5;
    const int bx = gx0 + lx, by = gy0 + ly;
    const bool inside = bx < fd.bxs && by < fd.bys;
    uint8_t a = 0;
    if (inside) a = acs[(size_t)by * fd.bxs + bx];
    const bool first = inside && (a & 0x80);
    const int s = a & 0x7f;
    const int n = first ? c_covered_x[s] * c_covered_y[s] : 1;
#pragma unroll
    for (int slot = 0; slot < 3; ++slot) {
      const int c = slot == 0 ? 1 : (slot == 1 ? 0 : 2);
      uint32_t v = 0;
      if (first) {
        const size_t bi = (size_t)c * nblk + (size_t)by * fd.bxs + bx;
        const int nz = nzcount[bi];
        v = 1 + (nz ? (uint32_t)(lastk[bi] - n + 1) : 0u);
      }
      cnt[i * 3 + slot] = v;
      sum += v;
    }
  }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t base = incl - sum;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
#pragma unroll
  for (int i = 0; i < 12; ++i) { s_off[t * 12 + i] = base; base += cnt[i]; }
  if (t == 255) { s_off[3072] = base; if (part == 0) token_counts[g] = base; }
  __syncthreads();
  // ---- emit: this CTA's slice of the entries; a warp takes 32 entries at a time, every lane
  // prefetching the metadata of one of them, then the warp walks the 32 entries together
  uint32_t* out = tokens + (size_t)g * kTokensPerGroupMax;
  // Rounds of 32 entries are drawn from a CTA-wide counter: the token count of a round varies by more than 10x with the
  // content, and a static split left the warps waiting 35 % of the kernel at the final barrier (ncu, profiles/r01g)
  constexpr int kPerCta = 3072 / kTokSplit;          // 768
  constexpr int kRounds = kPerCta / 32;              // 24
  for (;;) {
    int round = 0;
    if (lane == 0) round = (int)atomicAdd(&s_next, 1u);
    round = __shfl_sync(0xffffffffu, round, 0);
    if (round >= kRounds) break;
    const int e0 = part * kPerCta + round * 32;
    const int e = e0 + lane;
    uint32_t m_off = 0, m_count = 0, m_nztok = 0, m_misc = 0;   // misc: s | block_ctx << 8 | nz << 16
    {
      m_off = s_off[e];
      m_count = s_off[e + 1] - m_off;
      if (m_count) {
        const int blk = e / 3, slot = e - blk * 3;
        const int c = slot == 0 ? 1 : (slot == 1 ? 0 : 2);
        const int lx = blk & 31, ly = blk >> 5, bx = gx0 + lx, by = gy0 + ly;
        const size_t bi = (size_t)by * fd.bxs + bx;
        const int s = acs[bi] & 0x7f;
        const int block_ctx = c_block_ctx_map[(c < 2 ? c ^ 1 : 2) * kNumOrders + c_strategy_order[s]];
        const uint8_t* nzp = nzeros + (size_t)c * nblk;
        const int nz = nzcount[(size_t)c * nblk + bi];
        int pred;
        if (lx == 0) pred = ly == 0 ? 32 : nzp[bi - fd.bxs];
        else if (ly == 0) pred = nzp[bi - 1];
        else pred = (nzp[bi - fd.bxs] + nzp[bi - 1] + 1) >> 1;
        const int p = pred >= 64 ? 64 : pred;
        const int bucket = p < 8 ? p : 4 + (p >> 1);
        const uint32_t ctx = (uint32_t)(bucket * kNumBlockCtx + block_ctx);
        m_nztok = (ctx << 16) | (uint32_t)nz;
        m_misc = (uint32_t)s | ((uint32_t)block_ctx << 8) | ((uint32_t)nz << 16);
        out[m_off] = m_nztok;
        if (m_count > 1) {
          const size_t cblk = (size_t)g * kGroupBlocks + (size_t)ly * 32 + lx;
          const uint4* src = reinterpret_cast<const uint4*>(coeffs + (cblk * 3 + slot) * 64);
          uint4* dst = reinterpret_cast<uint4*>(&s_coef[warp][lane][0]);
          dst[0] = __ldg(src); dst[1] = __ldg(src + 1);
        }
        uint32_t tok, nb, bits;
        hybrid_encode((uint32_t)nz, tok, nb, bits);
        if (tok < 2) atomicAdd(&s_hot[ctx], tok ? 0x10000u : 1u); else atomicAdd(&hist[ctx * kAcAlphabet + tok], 1u);
      }
    }
    // coefficient tokens: a (block, channel) has ~4.5 of them on average, so four entries are walked side by side,
    // eight lanes each (the running non-zero count comes from the group's byte of the ballot)
    __syncwarp();
    // Entries with more than 32 coefficient tokens (the first blocks of 16/32/64-sized transforms carry up to 4095 of them,
    // their covered blocks none) are walked by the whole warp, one entry at a time, 32 scan positions per step: eight
    // lanes on such an entry kept the other warps of the CTA waiting at the final barrier (57 % of the kernel's stall
    // samples, profiles/r02k)
    unsigned big = __ballot_sync(0xffffffffu, m_count > 33);
    unsigned todo = __ballot_sync(0xffffffffu, m_count > 1) & ~big;
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t off = __shfl_sync(0xffffffffu, m_off, src);
      const uint32_t count = __shfl_sync(0xffffffffu, m_count, src);
      const uint32_t misc = __shfl_sync(0xffffffffu, m_misc, src);
      const int ee = e0 + src;
      const int blk = ee / 3, slot = ee - blk * 3;
      const int lx = blk & 31, ly = blk >> 5;
      const int s = misc & 0xff, block_ctx = (misc >> 8) & 0xff;
      int nz = (int)(misc >> 16);
      const int cx = c_covered_x[s], cy = c_covered_y[s], n = cx * cy, size = n * 64;
      const int log2n = 31 - __clz(n);
      const int histo_offset = kNumBlockCtx * kNonZeroBuckets + kZeroDensityContextCount * block_ctx;
      int prev_carry = nz > size / 16 ? 0 : 1;
      const int last = n + (int)count - 2;
      for (int k0 = n; k0 <= last; k0 += 32) {
        const int k = k0 + lane;
        int coef = 0;
        if (k <= last) {
          if (k < 16) {
            coef = s_coef[warp][src][k];
          } else {
            const int jj = k >> 6;
            const int jx = jj % cx, jy = jj / cx;
            const size_t cblk = (size_t)g * kGroupBlocks + (size_t)(ly + jy) * 32 + (lx + jx);
            coef = coeffs[(cblk * 3 + slot) * 64 + (k & 63)];
          }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, coef != 0);
        const int nz_here = nz - __popc(mask & ((1u << lane) - 1));
        const int prev = lane == 0 ? prev_carry : (int)((mask >> (lane - 1)) & 1);
        if (k <= last) {
          const int nzl = (nz_here + n - 1) >> log2n;
          const uint32_t ctx = (uint32_t)(histo_offset + (c_nnz_ctx[nzl] + c_freq_ctx[k >> log2n]) * 2 + prev);
          const uint32_t v = pack_signed(coef);
          out[off + 1 + (k - n)] = (ctx << 16) | v;
          uint32_t tok, nb, bits;
          hybrid_encode(v, tok, nb, bits);
          if (tok < 2) atomicAdd(&s_hot[ctx], tok ? 0x10000u : 1u); else atomicAdd(&hist[ctx * kAcAlphabet + tok], 1u);
        }
        nz -= __popc(mask);
        prev_carry = (int)(mask >> 31);
      }
    }
    const int gi = lane >> 3, gl = lane & 7;
    while (todo) {
      int j = -1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jj = todo ? __ffs(todo) - 1 : -1;
        if (todo) todo &= todo - 1;
        if (q == gi) j = jj;
      }
      const bool active = j >= 0;
      const int src = active ? j : 0;
      const uint32_t off = __shfl_sync(0xffffffffu, m_off, src);
      const uint32_t count = __shfl_sync(0xffffffffu, m_count, src);
      const uint32_t misc = __shfl_sync(0xffffffffu, m_misc, src);
      const int ee = e0 + src;
      const int blk = ee / 3, slot = ee - blk * 3;
      const int lx = blk & 31, ly = blk >> 5;
      const int s = misc & 0xff, block_ctx = (misc >> 8) & 0xff;
      int nz = (int)(misc >> 16);
      const int cx = c_covered_x[s], cy = c_covered_y[s], n = cx * cy, size = n * 64;
      const int log2n = 31 - __clz(n);
      const int histo_offset = kNumBlockCtx * kNonZeroBuckets + kZeroDensityContextCount * block_ctx;
      int prev_carry = nz > size / 16 ? 0 : 1;
      const int last = active ? n + (int)count - 2 : -1;  // scan position of the last non-zero coefficient
      for (int k0 = n; __any_sync(0xffffffffu, k0 <= last); k0 += 8) {
        const int k = k0 + gl;
        int coef = 0;
        if (k <= last) {
          if (k < 16) {
            coef = s_coef[warp][src][k];
          } else {
            const int jj = k >> 6;
            const int jx = jj % cx, jy = jj / cx;
            const size_t cblk = (size_t)g * kGroupBlocks + (size_t)(ly + jy) * 32 + (lx + jx);
            coef = coeffs[(cblk * 3 + slot) * 64 + (k & 63)];
          }
        }
        const unsigned mask = (__ballot_sync(0xffffffffu, coef != 0) >> (gi * 8)) & 0xFFu;
        const int nz_here = nz - __popc(mask & ((1u << gl) - 1));
        const int prev = gl == 0 ? prev_carry : (int)((mask >> (gl - 1)) & 1);
        if (k <= last) {
          const int nzl = (nz_here + n - 1) >> log2n;
          const uint32_t ctx = (uint32_t)(histo_offset + (c_nnz_ctx[nzl] + c_freq_ctx[k >> log2n]) * 2 + prev);
          const uint32_t v = pack_signed(coef);
          out[off + 1 + (k - n)] = (ctx << 16) | v;
          uint32_t tok, nb, bits;
          hybrid_encode(v, tok, nb, bits);
          if (tok < 2) atomicAdd(&s_hot[ctx], tok ? 0x10000u : 1u); else atomicAdd(&hist[ctx * kAcAlphabet + tok], 1u);
        }
        nz -= __popc(mask);
        prev_carry = (int)(mask >> 7);
      }
    }
    __syncwarp();   // the next round's lanes overwrite s_coef[warp]
  }
  __syncthreads();
  for (int i = t; i < kNumAcContexts; i += 256) {
    const uint32_t v = s_hot[i];
    if (v & 0xFFFFu) atomicAdd(&hist[i * kAcAlphabet], v & 0xFFFFu);
    if (v >> 16) atomicAdd(&hist[i * kAcAlphabet + 1], v >> 16);
  }
}

// ------------------------------------------------------------------------------------------ K9
constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = 512;
constexpr int kClusterWarps = kClusterCtas * kClusterThreads / 32;  // 128
constexpr int kCtxSlots = (kNumAcContexts + kClusterWarps * 32 - 1) / (kClusterWarps * 32);  // 2
constexpr int kCacheCap = 40;   // cached histogram rows per warp (16 warps x 40 x 256 B = 160 KB per CTA)

struct ClusterState {   // global scratch
  int assign[kNumAcContexts];
  uint32_t total[kNumAcContexts];
  int num_clusters;
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// candidate ordering for the argmax: larger key first, lower context index on ties
__device__ __forceinline__ bool better(long long ka, int ca, long long kb, int cb) { return ka > kb || (ka == kb && ca < cb); }

// Context ownership is static: warp w of the 128 owns contexts w, w + 128, w + 256, ...; lane i keeps
// (total, dist, assign) of the warp's i-th and (i + 32)-th context in registers for the whole
// kernel, so a round only loads the histogram rows themselves.
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads)
    k_cluster(const uint32_t* __restrict__ hist, const int* __restrict__ lut_g, ClusterState* __restrict__ st,
              uint8_t* __restrict__ cmap, uint32_t* __restrict__ cluster_hist, int max_clusters) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int s_lut[1025];
  __shared__ uint32_t s_hb[kAcAlphabet];
  __shared__ long long s_xb[kAcAlphabet];
  __shared__ long long s_cand_key[2][kClusterCtas];
  __shared__ int s_cand_ctx[2][kClusterCtas];
  __shared__ long long s_wkey[kClusterThreads / 32];
  __shared__ int s_wctx[kClusterThreads / 32];
  // per-warp cache of the histogram rows of the warp's non-empty contexts: every round re-reads them,
  // from shared memory (29 cycles) instead of L2 (~700 cycles, serialised by the reduction in between)
  extern __shared__ uint32_t s_cache[];   // [warps][kCacheCap][64]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int rank = (int)cluster.block_rank();
  const int gwarp = rank * (kClusterThreads / 32) + warp;
  uint32_t* my_cache = s_cache + (size_t)warp * kCacheCap * kAcAlphabet;
  for (int i = t; i < 1025; i += kClusterThreads) s_lut[i] = lut_g[i];
  __syncthreads();
  // ---- phase A: totals of the owned contexts
  uint32_t my_total[kCtxSlots];
  long long my_dist[kCtxSlots];
  int my_assign[kCtxSlots];
#pragma unroll
  for (int sl = 0; sl < kCtxSlots; ++sl) { my_total[sl] = 0; my_dist[sl] = -1; my_assign[sl] = 0; }
#pragma unroll
  for (int sl = 0; sl < kCtxSlots; ++sl) {
    for (int i0 = 0; i0 < 32; i0 += 4) {
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {       // four rows in flight
        const int c = gwarp + (sl * 32 + i0 + u) * kClusterWarps;
        v[u] = 0;
        if (c < kNumAcContexts) { const uint32_t* h = hist + (size_t)c * kAcAlphabet; v[u] = h[lane] + h[lane + 32]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v[u] += __shfl_xor_sync(0xffffffffu, v[u], d);
        if (lane == i0 + u) { my_total[sl] = v[u]; my_dist[sl] = v[u] ? 0x7fffffffffffffffll : -1; }
      }
    }
  }
  long long best_key = -1; int best_ctx = 0x7fffffff;
  unsigned ne_mask[kCtxSlots];
  int cache_base[kCtxSlots];
  {
    int acc = 0;
#pragma unroll
    for (int sl = 0; sl < kCtxSlots; ++sl) { ne_mask[sl] = __ballot_sync(0xffffffffu, my_total[sl] != 0); cache_base[sl] = acc; acc += __popc(ne_mask[sl]); }
  }
#pragma unroll
  for (int sl = 0; sl < kCtxSlots; ++sl) {
    const int c = gwarp + (sl * 32 + lane) * kClusterWarps;
    if (my_total[sl] && better((long long)my_total[sl], c, best_key, best_ctx)) { best_key = (long long)my_total[sl]; best_ctx = c; }
    if (c < kNumAcContexts) st->total[c] = my_total[sl];
    unsigned mask = ne_mask[sl];
    while (mask) {                      // copy the non-empty rows into the cache
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const int r = cache_base[sl] + __popc(ne_mask[sl] & ((1u << i) - 1));
      if (r < kCacheCap) {
        const uint32_t* h = hist + (size_t)(gwarp + (sl * 32 + i) * kClusterWarps) * kAcAlphabet;
        my_cache[r * kAcAlphabet + lane] = h[lane];
        my_cache[r * kAcAlphabet + lane + 32] = h[lane + 32];
      }
    }
  }
  __syncwarp();
  int K = 0;
  int parity = 0;
  for (;;) {
    // ---- cluster-wide argmax of (best_key, best_ctx): lanes -> warp -> CTA -> cluster (DSMEM)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const long long ok = __shfl_xor_sync(0xffffffffu, best_key, d);
      const int oc = __shfl_xor_sync(0xffffffffu, best_ctx, d);
      if (better(ok, oc, best_key, best_ctx)) { best_key = ok; best_ctx = oc; }
    }
    if (lane == 0) { s_wkey[warp] = best_key; s_wctx[warp] = best_ctx; }
    __syncthreads();
    if (warp == 0) {
      // the CTA's best of its 16 warps (a butterfly over lanes 0..15), then lane r tells rank r
      long long k = lane < kClusterThreads / 32 ? s_wkey[lane] : -1;
      int c = lane < kClusterThreads / 32 ? s_wctx[lane] : 0x7fffffff;
#pragma unroll
      for (int d = 8; d >= 1; d >>= 1) {
        const long long ok = __shfl_xor_sync(0xffffffffu, k, d);
        const int oc = __shfl_xor_sync(0xffffffffu, c, d);
        if (better(ok, oc, k, c)) { k = ok; c = oc; }
      }
      if (lane < kClusterCtas) {
        long long* rk = cluster.map_shared_rank(&s_cand_key[parity][0], lane);
        int* rc = cluster.map_shared_rank(&s_cand_ctx[parity][0], lane);
        rk[rank] = k; rc[rank] = c;
      }
    }
    cluster.sync();
    long long sk = s_cand_key[parity][0]; int seed = s_cand_ctx[parity][0];
    for (int r = 1; r < kClusterCtas; ++r)
      if (better(s_cand_key[parity][r], s_cand_ctx[parity][r], sk, seed)) { sk = s_cand_key[parity][r]; seed = s_cand_ctx[parity][r]; }
    parity ^= 1;
    if (sk < 0) break;                                        // no non-empty context at all
    if (K > 0 && sk < ((long long)64 << 20)) break;           // nothing far enough from every seed
    const int k = K++;
    // ---- distances to the new seed
    if (t < kAcAlphabet) {
      const uint32_t b = hist[(size_t)seed * kAcAlphabet + t];
      s_hb[t] = b;
      s_xb[t] = xlogx(b, s_lut);
    }
    __syncthreads();
    uint32_t tb = 0;
    for (int s = 0; s < kAcAlphabet; ++s) tb += s_hb[s];      // 64 broadcast LDS: cheaper than another barrier
    const long long xtb = xlogx(tb, s_lut);
    best_key = -1; best_ctx = 0x7fffffff;
#pragma unroll
    for (int sl = 0; sl < kCtxSlots; ++sl) {
      // four contexts per trip: their partial sums and the four 64-bit butterflies are independent, so the shuffle
      // latencies of one overlap the arithmetic of the others (integer sums: any order gives the same value)
      unsigned mask = ne_mask[sl];
      const uint32_t b0 = s_hb[lane], b1 = s_hb[lane + 32];
      const long long xb0 = s_xb[lane], xb1 = s_xb[lane + 32];
      while (mask) {
        int idx[4];
        long long acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          idx[u] = mask ? __ffs(mask) - 1 : -1;
          if (mask) mask &= mask - 1;
          acc[u] = 0;
          if (idx[u] >= 0) {
            const int c = gwarp + (sl * 32 + idx[u]) * kClusterWarps;
            const int r = cache_base[sl] + __popc(ne_mask[sl] & ((1u << idx[u]) - 1));
            const uint32_t* h = r < kCacheCap ? my_cache + r * kAcAlphabet : hist + (size_t)c * kAcAlphabet;
            const uint32_t a0 = h[lane], a1 = h[lane + 32];
            if (a0 && b0) acc[u] += xlogx(a0 + b0, s_lut) - xlogx(a0, s_lut) - xb0;
            if (a1 && b1) acc[u] += xlogx(a1 + b1, s_lut) - xlogx(a1, s_lut) - xb1;
          }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], d);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (idx[u] >= 0 && lane == idx[u]) {
            const int c = gwarp + (sl * 32 + idx[u]) * kClusterWarps;
            const uint32_t ta = my_total[sl];
            long long d = xlogx(ta + tb, s_lut) - xlogx(ta, s_lut) - xtb - acc[u];
            if (c == seed) d = 0;
            if (d < my_dist[sl]) { my_dist[sl] = d; my_assign[sl] = k; }
          }
        }
      }
      const int c = gwarp + (sl * 32 + lane) * kClusterWarps;
      if (my_total[sl] && better(my_dist[sl], c, best_key, best_ctx)) { best_key = my_dist[sl]; best_ctx = c; }
    }
    __syncthreads();  // s_hb / s_xb are rewritten next round
    if (K >= max_clusters) break;
  }
  if (K == 0) K = 1;
#pragma unroll
  for (int sl = 0; sl < kCtxSlots; ++sl) {
    const int c = gwarp + (sl * 32 + lane) * kClusterWarps;
    if (c < kNumAcContexts) st->assign[c] = my_assign[sl];
  }
  // ---- cluster histograms (sum of members)
#pragma unroll
  for (int sl = 0; sl < kCtxSlots; ++sl) {
    unsigned mask = ne_mask[sl];
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const int c = gwarp + (sl * 32 + i) * kClusterWarps;
      const int k = __shfl_sync(0xffffffffu, my_assign[sl], i);
      const int r = cache_base[sl] + __popc(ne_mask[sl] & ((1u << i) - 1));
      const uint32_t* h = r < kCacheCap ? my_cache + r * kAcAlphabet : hist + (size_t)c * kAcAlphabet;
      for (int s = lane; s < kAcAlphabet; s += 32) { const uint32_t v = h[s]; if (v) atomicAdd(&cluster_hist[k * kAcAlphabet + s], v); }
    }
  }
  cluster.sync();
  if (rank == 0) {
    // context map: empty contexts inherit the cluster of the previous non-empty context (0 before the first)
    __shared__ int s_carry[kClusterThreads];
    constexpr int kChunk = (kNumAcContexts + kClusterThreads - 1) / kClusterThreads;
    const int c0 = t * kChunk, c1 = min(c0 + kChunk, kNumAcContexts);
    int lastv = -1;
    for (int c = c0; c < c1; ++c) if (st->total[c]) lastv = st->assign[c];
    s_carry[t] = lastv;
    __syncthreads();
    if (t == 0) {
      int run = 0;
      for (int i = 0; i < kClusterThreads; ++i) { const int v = s_carry[i]; s_carry[i] = run; if (v >= 0) run = v; }
      st->num_clusters = K;
    }
    __syncthreads();
    int prev = s_carry[t];
    for (int c = c0; c < c1; ++c) { if (st->total[c]) prev = st->assign[c]; cmap[c] = (uint8_t)prev; }
  }
}

// ------------------------------------------------------------------------------------------ K9b
__global__ void __launch_bounds__(32) k_ans_tables(const uint32_t* __restrict__ cluster_hist,
                                                   const ClusterState* __restrict__ st, uint16_t* __restrict__ norm,
                                                   uint16_t* __restrict__ rmap, AnsSymInfo* __restrict__ info,
                                                   uint32_t* __restrict__ hdr_bits, uint32_t* __restrict__ hdr_len) {
  __shared__ uint16_t s_norm[kAcAlphabet];
  __shared__ uint16_t s_scratch[1024];
  const int k = blockIdx.x, lane = threadIdx.x;
  if (k >= st->num_clusters) return;
  if (lane == 0) {
    normalize_counts(cluster_hist + (size_t)k * kAcAlphabet, kAcAlphabet, s_norm);
    BitWriterDev w;
    w.init(hdr_bits + (size_t)k * 64);
    write_ans_histogram(s_norm, kAcAlphabet, w);
    w.flush();
    hdr_len[k] = w.bits();
  }
  __syncwarp();
  for (int s = lane; s < kAcAlphabet; s += 32) norm[(size_t)k * kAcAlphabet + s] = s_norm[s];
  build_reverse_map(s_norm, kAcAlphabet, s_scratch, rmap + (size_t)k * kAnsTabSize, info + (size_t)k * kAcAlphabet, lane);
}

// ------------------------------------------------------------------------------------------ K10
// One warp per AC group, kAnsWarps groups per CTA.  The CTA first stages EVERY table the chain
// touches in shared memory (reverse maps of all clusters, symbol info, context map: <= 212 KB),
// so the state-dependent lookup of each rANS step is an LDS (29 cycles) instead of an L1/L2 miss.
// Per 32 tokens: lanes fetch + classify one token each (the next chunk's tokens are prefetched
// while the current chain runs); the serial chain runs over shuffles and only computes the
// state, each lane capturing its own token's renormalisation word; then the 32 variable-length
// pieces are placed by a warp prefix sum and ORed into a 34-word shared staging window that is
// stored with coalesced 32-bit writes.  The stream is produced back to front and ENDS at word
// kTokensPerGroupMax of the group's arena; start_bit[g] = position of its first bit.
constexpr int kStageWords = 34;

template <int kAnsWarps>
__global__ void __launch_bounds__(kAnsWarps * 32) k_ans_groups(const uint32_t* __restrict__ tokens,
                                                               const uint32_t* __restrict__ token_counts,
                                                               const uint8_t* __restrict__ cmap_g,
                                                               const AnsSymInfo* __restrict__ info_g,
                                                               const uint16_t* __restrict__ rmap_g,
                                                               const int* __restrict__ num_clusters_p, int num_groups,
                                                               uint32_t* __restrict__ work_counter,
                                                               uint32_t* __restrict__ out_arena,
                                                               unsigned long long* __restrict__ start_bit) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int K = *num_clusters_p;
  uint16_t* s_rmap = reinterpret_cast<uint16_t*>(smem);                                  // [K][4096]
  // (the per-symbol info is only read by the lane-parallel operand preparation, which runs a chunk ahead of the chain:
  // it stays in global memory / L1, and its 12 KB of shared memory hold 16-byte chain operands instead)
  uint8_t* s_cmap = smem + (size_t)K * kAnsTabSize * 2;                                    // [7425 -> 7440]
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_cmap + 7440);                          // [warps][34]
  uint4* s_ops = reinterpret_cast<uint4*>(s_stage + kAnsWarps * kStageWords);                // [warps][2][32]
  uint32_t* s_states = reinterpret_cast<uint32_t*>(s_ops + kAnsWarps * 64);                  // [warps][32]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  {
    // The CTA's tables — K reverse maps of 8 KB and the context map, up to 199 KB, contiguous in global memory — are
    // brought in by the copy engine: one thread issues bulk asynchronous copies (cp.async.bulk, 32 KB each) that complete
    // on an mbarrier, every thread waits on its phase.  (The per-thread LDG -> STS loop this replaces took 48 round
    // trips to L2 per thread before any chain could start.)
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (t == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
      const uint32_t bytes_r = (uint32_t)K * kAnsTabSize * 2, bytes_c = 7440;   // (the context map buffer is padded past 7425)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes_r + bytes_c) : "memory");
      const uint32_t dst_r = (uint32_t)__cvta_generic_to_shared(s_rmap);
      for (uint32_t off = 0; off < bytes_r; off += 32768u) {
        const uint32_t n = min(32768u, bytes_r - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_r + off), "l"(reinterpret_cast<const uint8_t*>(rmap_g) + off), "r"(n), "r"(mbar) : "memory");
      }
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(s_cmap)), "l"(cmap_g), "r"(bytes_c), "r"(mbar) : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ANS_TABLES_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra ANS_TABLES_DONE;\n"
        "bra ANS_TABLES_WAIT;\n"
        "ANS_TABLES_DONE:\n"
        "}\n" ::"r"(mbar) : "memory");
  }
  __syncthreads();
  const uint32_t rmap_saddr = (uint32_t)__cvta_generic_to_shared(s_rmap);
  // warps take groups from a shared counter (longest-processing-time order is not needed: a warp that
  // drew a short group simply draws again), so a CTA stays resident only as long as it has work
  for (;;) {
  int g = 0;
  if (lane == 0) g = (int)atomicAdd(work_counter, 1u);
  g = __shfl_sync(0xffffffffu, g, 0);
  if (g >= num_groups) return;
  const uint32_t* tk = tokens + (size_t)g * kTokensPerGroupMax;
  uint32_t* out = out_arena + (size_t)g * kTokensPerGroupMax;
  uint32_t* stage = s_stage + warp * kStageWords;
  const int n = (int)token_counts[g];
  uint32_t state = kAnsInitState;
  long long end_bit = (long long)kTokensPerGroupMax * 32;  // stream position where the next (earlier) piece ends
  uint32_t carry = 0;                                      // bits of the partially filled word containing end_bit
  uint4* ops = s_ops + warp * 64;                          // two slots of 32 per-token chain operands (thr, -freq, rcp, table)
  uint32_t* states = s_states + warp * 32;                 // the state every step of the current chunk started from

  // per-token operands of the chain — renormalisation threshold (freq << 20) - 1, -freq, reciprocal, shared byte address
  // of the symbol's reverse-map run — and the token's extra bits
  auto prep = [&](uint32_t tkn, bool valid, uint4& op, uint32_t& nb, uint32_t& bits) {
    op = make_uint4(0xFFFFFFFFu, 0u - 4096u, 0x00100000u, rmap_saddr); nb = 0; bits = 0;
    if (valid) {
      const uint32_t cl = s_cmap[tkn >> 16];
      uint32_t tok;
      hybrid_encode(tkn & 0xFFFF, tok, nb, bits);
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(info_g + cl * kAcAlphabet + tok));
      const uint32_t f = raw.x & 0xFFFFu, base = raw.x >> 16;
      op = make_uint4((f << 20) - 1, 0u - f, raw.y, rmap_saddr + 2u * (cl * kAnsTabSize + base));
    }
  };
  // one chain step on operands op = (thr, -freq, rcp, table); lane j records what step j pushed out
  auto step = [&](uint4 op, int j, uint32_t& my_o16, int& my_emit) {
    const uint32_t thr = op.x, negf = op.y, rc = op.z, tab_a = op.w;
    const uint32_t f = 0u - negf, tab_b = tab_a + 2 * negf;
    const bool emit = state > thr;
    if (lane == j) { my_o16 = state & 0xFFFF; my_emit = emit; }
    const uint32_t x2 = emit ? (state >> 16) : state;
    const uint32_t q = __umulhi(x2, rc);
    const uint32_t r = q * negf + x2;
    const bool fix = r >= f;
    const uint32_t a_lo = tab_a + 2 * r, a_hi = tab_b + 2 * r;
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(fix ? a_hi : a_lo));
    state = ((fix ? q + 1 : q) << kAnsLogTabSize) + v;
  };
  // the same step for the pipelined chunks: instead of every lane testing `lane == j` (a compare and two selects per
  // step), lane 0 stores the state the step started from; lane j reads states[j] after the chain and derives what its
  // token pushed out from its own copy of the operands
  auto step_rec = [&](uint4 op, int j) {
    const uint32_t thr = op.x, negf = op.y, rc = op.z, tab_a = op.w;
    const uint32_t f = 0u - negf, tab_b = tab_a + 2 * negf;
    const bool emit = state > thr;
    if (lane == 0) states[j] = state;
    const uint32_t x2 = emit ? (state >> 16) : state;
    const uint32_t q = __umulhi(x2, rc);
    const uint32_t r = q * negf + x2;
    const bool fix = r >= f;
    const uint32_t a_lo = tab_a + 2 * r, a_hi = tab_b + 2 * r;
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(fix ? a_hi : a_lo));
    state = ((fix ? q + 1 : q) << kAnsLogTabSize) + v;
  };
  // Placing a chunk's pieces (lane j's piece = [renorm word (16)][extra bits (nb)], earlier lanes later in the
  // stream) in three phases separated by warp barriers, so that the chain of the NEXT chunk can run between them.
  struct PackState { int len, incl, nwords, first_final; uint32_t val; long long lo_bit, base_word; };
  auto pack_a = [&](uint32_t nb, uint32_t bits, uint32_t o16, int emit, bool valid, PackState& P) {
    P.len = valid ? (int)nb + (emit ? 16 : 0) : 0;
    P.val = emit ? (o16 | (bits << 16)) : bits;
    int incl = P.len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    P.incl = incl;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    P.lo_bit = end_bit - total;                               // first stream bit of this chunk
    P.base_word = P.lo_bit >> 5;                              // lowest word touched
    const long long top_word = end_bit >> 5;                  // word holding `carry` (if end_bit is unaligned)
    P.nwords = (int)(((end_bit + 31) >> 5) - P.base_word);    // words covering [lo_bit, end_bit): <= 31
    stage[lane] = (lane == (int)(top_word - P.base_word) && (end_bit & 31)) ? carry : 0;
    if (lane < 2) stage[32 + lane] = (32 + lane == (int)(top_word - P.base_word) && (end_bit & 31)) ? carry : 0;
  };
  auto pack_b = [&](const PackState& P) {
    if (P.len) {
      const long long p = end_bit - P.incl;                   // stream position of my piece
      const int off = (int)(p - (P.base_word << 5));
      const unsigned long long v = (unsigned long long)P.val << (off & 31);
      atomicOr(&stage[off >> 5], (uint32_t)v);
      if ((off & 31) + P.len > 32) atomicOr(&stage[(off >> 5) + 1], (uint32_t)(v >> 32));
    }
  };
  auto pack_c = [&](const PackState& P) {
    // words strictly above the (possibly partial) lowest word are final
    const int first_final = (P.lo_bit & 31) ? 1 : 0;
    for (int w = first_final + lane; w < P.nwords; w += 32) {
      if (P.base_word + w < (long long)kTokensPerGroupMax) out[P.base_word + w] = stage[w];
    }
    carry = stage[0];
    end_bit = P.lo_bit;
  };

  // ---- the stream tail: the (possibly partial) last chunk of tokens, not pipelined
  const int m0 = n > 0 ? ((n - 1) & 31) + 1 : 0;
  uint32_t nb_p = 0, bits_p = 0, o16_p = 0; int emit_p = 0;   // pieces of the chunk whose chain has run, not yet placed
  bool valid_p = false;
  if (n > 0) {
    uint4 op;
    const uint32_t tkn = lane < m0 ? tk[n - 1 - lane] : 0;
    prep(tkn, lane < m0, op, nb_p, bits_p);
    for (int j = 0; j < m0; ++j) {
      uint4 o;
      o.x = __shfl_sync(0xffffffffu, op.x, j); o.y = __shfl_sync(0xffffffffu, op.y, j);
      o.z = __shfl_sync(0xffffffffu, op.z, j); o.w = __shfl_sync(0xffffffffu, op.w, j);
      step(o, j, o16_p, emit_p);
    }
    valid_p = lane < m0;
  }
  // ---- full chunks, software-pipelined: while the serial chain of chunk c runs (a dependent sequence that leaves
  // most issue slots empty), the same warp prepares the operands of chunk c + 1 and places the pieces of chunk c - 1
  const int full = n > 0 ? (n - m0) >> 5 : 0;
  uint32_t nb_c = 0, bits_c = 0, thr_c = 0xFFFFFFFFu;   // this lane's token of the chunk whose chain runs next
  if (full > 0) {
    uint4 op;
    prep(tk[n - m0 - 1 - lane], true, op, nb_c, bits_c);
    ops[lane] = op;
    thr_c = op.x;
  }
  uint32_t tok_next = full > 1 ? tk[n - m0 - 32 - 1 - lane] : 0;
  for (int c = 0; c < full; ++c) {
    const uint4* cur = ops + (c & 1) * 32;
    uint4* nxt = ops + ((c + 1) & 1) * 32;
    __syncwarp();   // operands of chunk c visible; stage free
    uint32_t o16_c = 0; int emit_c = 0;
    PackState P;
    // (a) first quarter of the chain || operands of the next chunk
    uint32_t nb_n = 0, bits_n = 0, thr_n = 0xFFFFFFFFu;
    {
      uint4 op;
      prep(tok_next, c + 1 < full, op, nb_n, bits_n);
      nxt[lane] = op;
      thr_n = op.x;
      const int in = n - m0 - 32 * (c + 2) - 1 - lane;
      tok_next = (c + 2 < full) ? tk[in] : 0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { step_rec(cur[j], j); }
    // (b) second quarter || prefix sum of the previous chunk's piece lengths, staging window cleared
    pack_a(nb_p, bits_p, o16_p, emit_p, valid_p, P);
#pragma unroll
    for (int j = 8; j < 16; ++j) { step_rec(cur[j], j); }
    __syncwarp();
    // (c) third quarter || the previous chunk's pieces OR-ed into the window
    pack_b(P);
#pragma unroll
    for (int j = 16; j < 24; ++j) { step_rec(cur[j], j); }
    __syncwarp();
    // (d) last quarter || finished words stored
    pack_c(P);
#pragma unroll
    for (int j = 24; j < 32; ++j) { step_rec(cur[j], j); }
    __syncwarp();
    {
      const uint32_t my_state = states[lane];
      o16_c = my_state & 0xFFFF;
      emit_c = my_state > thr_c ? 1 : 0;
    }
    nb_p = nb_c; bits_p = bits_c; o16_p = o16_c; emit_p = emit_c; valid_p = true;
    nb_c = nb_n; bits_c = bits_n; thr_c = thr_n;
  }
  // ---- drain: the pieces of the last chunk whose chain has run
  if (n > 0) {
    PackState P;
    __syncwarp();
    pack_a(nb_p, bits_p, o16_p, emit_p, valid_p, P);
    __syncwarp();
    pack_b(P);
    __syncwarp();
    pack_c(P);
    __syncwarp();
  }
  // final state (32 bits) in front, then flush the partial word
  {
    const long long lo_bit = end_bit - 32;
    const long long base_word = lo_bit >> 5;
    if (lane == 0) {
      const int sh = (int)(lo_bit & 31);
      if (sh == 0) out[base_word] = state;
      else {
        out[base_word + 1] = (carry & ~((1u << sh) - 1)) | (state >> (32 - sh));   // carry holds bits >= sh of that word
        out[base_word] = state << sh;
      }
      start_bit[g] = (unsigned long long)lo_bit;
    }
  }
  __syncwarp();
  }  // next group
}

// ------------------------------------------------------------------------------------------ launchers
size_t cluster_state_bytes() { return sizeof(ClusterState); }

void launch_tokenize(const uint8_t* acs, const uint8_t* nzeros, const uint16_t* nzcount, const uint16_t* lastk,
                     const int16_t* coeffs, const FrameDim& fd, uint32_t* tokens, uint32_t* token_counts, uint32_t* hist,
                     cudaStream_t s) {
  // (function attributes are per device: set on every launch, a context may live on any GPU of the process)
  cudaFuncSetAttribute(k_tokenize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kNumAcContexts * sizeof(uint32_t)));
  ++g_kernel_launches;
  k_tokenize<<<dim3(fd.num_groups, kTokSplit), 256, kNumAcContexts * sizeof(uint32_t), s>>>(acs, nzeros, nzcount, lastk, coeffs, fd, tokens, token_counts, hist);
}

void launch_cluster(const uint32_t* hist, const int* lut, void* state, uint8_t* cmap, uint32_t* cluster_hist,
                    cudaStream_t s) {
  ++g_kernel_launches;
  const size_t smem = (size_t)(kClusterThreads / 32) * kCacheCap * kAcAlphabet * sizeof(uint32_t);
  // (function attributes are per device: set on every launch, a context may live on any GPU of the process)
  cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_cluster<<<kClusterCtas, kClusterThreads, smem, s>>>(hist, lut, (ClusterState*)state, cmap, cluster_hist, kMaxClusters);
}

void launch_ans_tables(const uint32_t* cluster_hist, const void* state, uint16_t* norm, uint16_t* rmap, void* info,
                       uint32_t* hdr_bits, uint32_t* hdr_len, cudaStream_t s) {
  ++g_kernel_launches;
  k_ans_tables<<<kMaxClusters, 32, 0, s>>>(cluster_hist, (const ClusterState*)state, norm, rmap, (AnsSymInfo*)info,
                                           hdr_bits, hdr_len);
}

template <int kAnsWarps>
static void launch_ans_groups_w(const uint32_t* tokens, const uint32_t* token_counts, const uint8_t* cmap, const void* info,
                                const uint16_t* rmap, const int* num_clusters, uint32_t* work_counter, int groups_per_warp,
                                uint32_t* out_arena, unsigned long long* start_bit, int num_groups, cudaStream_t s) {
  const size_t smem = (size_t)kMaxClusters * kAnsTabSize * 2 + 7440 + kAnsWarps * kStageWords * 4 + kAnsWarps * 64 * sizeof(uint4) +
                      kAnsWarps * 32 * 4;
  cudaFuncSetAttribute(k_ans_groups<kAnsWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int per_cta = kAnsWarps * (groups_per_warp < 1 ? 1 : groups_per_warp);
  cudaMemsetAsync(work_counter, 0, 4, s);
  k_ans_groups<kAnsWarps><<<(num_groups + per_cta - 1) / per_cta, kAnsWarps * 32, smem, s>>>(
      tokens, token_counts, cmap, (const AnsSymInfo*)info, rmap, num_clusters, num_groups, work_counter, out_arena, start_bit);
}

// warps_per_cta: 8 for a single image (135 chains of a 4K frame spread over 17 SMs, two warps per scheduler: a chain
// issues ~25 % of the cycles, so four of them on one scheduler slow each other down), 16 in batch mode (a CTA holds the
// whole SM's shared memory for its reverse maps: more chains per CTA = less SM time taken from the other pipelines)
void launch_ans_groups(const uint32_t* tokens, const uint32_t* token_counts, const uint8_t* cmap, const void* info,
                       const uint16_t* rmap, const int* num_clusters, uint32_t* work_counter, int groups_per_warp,
                       int warps_per_cta, uint32_t* out_arena, unsigned long long* start_bit, int num_groups, cudaStream_t s) {
  ++g_kernel_launches;
  if (warps_per_cta >= 16)
    launch_ans_groups_w<16>(tokens, token_counts, cmap, info, rmap, num_clusters, work_counter, groups_per_warp, out_arena,
                            start_bit, num_groups, s);
  else if (warps_per_cta >= 8)
    launch_ans_groups_w<8>(tokens, token_counts, cmap, info, rmap, num_clusters, work_counter, groups_per_warp, out_arena,
                           start_bit, num_groups, s);
  else
    launch_ans_groups_w<4>(tokens, token_counts, cmap, info, rmap, num_clusters, work_counter, groups_per_warp, out_arena,
                           start_bit, num_groups, s);
}

int cluster_num_clusters_offset() { return (int)offsetof(ClusterState, num_clusters); }

}  // namespace jxlb
