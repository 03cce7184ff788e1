// tcgen05 / TMEM / TMA tiles for the KGEB_DOT all-entity forms (DistMult, ComplEx, CP, SimplE, RESCAL):
//   S[128 x 128] = A[128 x d] * B[128 x d]^T,  fp32 operands read in place as TF32 (kind::tf32), fp32
//   accumulation in tensor memory.  Both operands are K-major 128-byte-swizzled slabs written by TMA
//   straight from the row-major fp32 tables (no conversion pass, no re-layout of the table).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer   (query blocks -> resident smem, table slabs -> ring of stages)
//   warp 1    MMA issuer     (one thread issues tcgen05.mma; tcgen05.commit frees smem / publishes TMEM)
//   warp 2    TMEM allocator
//   warps 4-7 epilogue       (tcgen05.ld -> registers -> fused epilogue; TMEM double-buffered)
// A job = (group of NQB query blocks, contiguous chunk of entity tiles): the query blocks stay resident
// in shared memory for the whole chunk, only table slabs stream.  Epilogue modes:
//   MODE_SCORES  A = table tile, B = query block : lane = entity  -> coalesced stores of X[q, e]
//   MODE_STATS   A = query block, B = table tile : lane = query   -> online max/sum-exp | softplus sums
//   MODE_RANK    A = query block, B = table tile : lane = query   -> rank / tie counts with CSR filters
#include <climits>
#include "tc_common.cuh"

namespace kgeb {
namespace tc {

constexpr int TILE = 128;                 // rows per operand tile (UMMA M = N = 128)
constexpr int SLAB_BYTES = TILE * 128;    // 16 KiB: 128 rows x 128 B
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int EPQ = 4;                              // epilogue warps per TMEM lane quadrant (latency hiding)
constexpr int NUM_EPI_THREADS = 4 * EPQ * 32;       // 512
constexpr int NUM_THREADS = 128 + NUM_EPI_THREADS;  // warps 0-3: TMA / MMA / TMEM alloc / idle; 4-19: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int COLS_PER_WARP = TILE / EPQ;           // 32 accumulator columns per epilogue warp
constexpr int COMB_BYTES = EPQ * 2 * TILE * 3 * 4;  // per-job row-statistics exchange between column parts

enum { MODE_SCORES = 0, MODE_STATS = 1, MODE_RANK = 2 };

struct Params {
  int64_t B;          // query rows
  int64_t n_ent;      // table rows of this shard
  int64_t e_lo;       // global id of table row 0
  int d;
  int ks;             // K slabs = ceil(d / 32)
  int stages;         // ring depth
  int64_t n_qgroups;  // ceil(ceil(B/128) / NQB)
  int64_t n_tiles;    // ceil(n_ent / 128)
  int64_t chunks;     // entity chunks
  int64_t tiles_per_chunk;
  // MODE_SCORES
  float* out; int64_t ld; int64_t col_off;
  // MODE_STATS
  int loss; float offset; float* partial;  // [chunks][B][4]
  // MODE_RANK
  const float* true_score; const void* true_ent; int idx64;
  const int64_t* f_off; const int64_t* f_col; const int64_t* t_off; const int64_t* t_col;
  unsigned long long* counts;
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int MODE, int NQB, bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_tiles_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KS = p.ks, STAGES = p.stages;
  constexpr int SLAB_K = Elem<BF16>::kSlabK, UMMA_K = Elem<BF16>::kUmmaK;
  uint8_t* q_smem = smem;                                  // [NQB][KS] slabs, resident per job
  uint8_t* ring = smem + (size_t)NQB * KS * SLAB_BYTES;    // [STAGES] slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * SLAB_BYTES);
  uint64_t* full = bars;                 // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;   // [MAX_STAGES]
  uint64_t* q_full = bars + 2 * MAX_STAGES;
  uint64_t* q_empty = q_full + 1;
  uint64_t* t_full = q_full + 2;         // [2]
  uint64_t* t_empty = q_full + 4;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 6);
  float* comb = reinterpret_cast<float*>(q_full + 8);  // [EPQ][NQB][TILE][3]
  constexpr int TMEM_COLS = 2 * NQB * TILE;  // double-buffered accumulators (256 or 512 columns)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], NUM_EPI_THREADS / 32);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_qblocks = (p.B + TILE - 1) / TILE;
  const int64_t n_jobs = p.n_qgroups * p.chunks;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {   // single elected thread: see the MMA issuer
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const int64_t qg = job % p.n_qgroups, ch = job / p.n_qgroups;
        const int nqb = (int)min((int64_t)NQB, n_qblocks - qg * NQB);
        mbar_wait(q_empty, qphase ^ 1);
        mbar_expect_tx(q_full, (uint32_t)(nqb * KS * SLAB_BYTES));
        for (int qb = 0; qb < nqb; ++qb)
          for (int k = 0; k < KS; ++k)
            tma_load_2d(q_smem + (size_t)(qb * KS + k) * SLAB_BYTES, &tm_q, k * SLAB_K,
                        (int32_t)((qg * NQB + qb) * TILE), q_full);
        qphase ^= 1;
        const int64_t t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
        for (int64_t t = t0; t < t1; ++t)
          for (int k = 0; k < KS; ++k) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], SLAB_BYTES);
            tma_load_2d(ring + (size_t)stage * SLAB_BYTES, &tm_w, k * SLAB_K, (int32_t)(t * TILE), &full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // under one elect.sync predicate the compiler keeps the descriptors in uniform registers (tc_bwd.cu, MMA1 issuer)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(TILE, TILE, 0, 0, Elem<BF16>::kFmt);
      int stage = 0, buf = 0;
      uint32_t phase = 0, qphase = 0, tphase[2] = {0, 0};
      for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const int64_t qg = job % p.n_qgroups, ch = job / p.n_qgroups;
        const int nqb = (int)min((int64_t)NQB, n_qblocks - qg * NQB);
        mbar_wait(q_full, qphase);
        qphase ^= 1;
        tc_fence_after();
        const int64_t t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
        for (int64_t t = t0; t < t1; ++t) {
          mbar_wait(&t_empty[buf], tphase[buf] ^ 1);
          tc_fence_after();
          for (int k = 0; k < KS; ++k) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(ring + (size_t)stage * SLAB_BYTES);
            for (int qb = 0; qb < nqb; ++qb) {
              const uint32_t q_addr = smem_u32(q_smem + (size_t)(qb * KS + k) * SLAB_BYTES);
              const uint32_t acc = tmem_base + (uint32_t)((buf * NQB + qb) * TILE);
#pragma unroll
              for (int kk = 0; kk < SLAB_K / UMMA_K; ++kk) {
                const uint64_t qd = make_desc(q_addr + kk * 32, 16, 1024);  // UMMA_K elements = 32 bytes
                const uint64_t wd = make_desc(w_addr + kk * 32, 16, 1024);
                if (MODE == MODE_SCORES) umma<BF16>(acc, wd, qd, idesc, (k | kk) != 0);  // lanes = entities
                else                     umma<BF16>(acc, qd, wd, idesc, (k | kk) != 0);  // lanes = queries
              }
            }
            umma_commit(&empty[stage]);  // frees the slab once these MMAs have read it
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&t_full[buf]);     // accumulators of this tile are complete
          tphase[buf] ^= 1;
          buf ^= 1;
        }
        umma_commit(q_empty);            // resident query blocks may be overwritten
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue (16 warps: lane quadrant x column part) ================================
    const int ew = warp - EPI_WARP0;
    const int quad = ew & 3;                    // == warp % 4 : TMEM lane quadrant this warp may access
    const int part = ew >> 2;                   // which 32 accumulator columns of the 128
    const int trow = quad * 32 + lane;          // row of the A tile owned by this thread
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int cw0 = part * COLS_PER_WARP;       // first column of this warp
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const int64_t qg = job % p.n_qgroups, ch = job / p.n_qgroups;
      const int nqb = (int)min((int64_t)NQB, n_qblocks - qg * NQB);
      const int64_t t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);

      // per-job running state (MODE_STATS / MODE_RANK: thread = query row trow of block qb, columns of its part)
      float s0[NQB], s1[NQB], s2[NQB];
      int cnt[NQB][6];
      float tsc[NQB];
      // MODE_RANK: entity columns are tracked relative to the first entity of the chunk (32-bit); the next unconsumed
      // filter column of each row sits in a register so that the per-tile "any filter entry here?" test costs no load
      int trel[NQB], fcur[NQB], fend[NQB], fnxt[NQB], ucur[NQB], uend[NQB], unxt[NQB];
      const int64_t chunk_ent0 = p.e_lo + t0 * TILE;
      auto rel_of = [&](const int64_t* col, int cur, int end) -> int {
        if (cur >= end) return INT_MAX;
        const int64_t r = col[cur] - chunk_ent0;
        return r >= (int64_t)INT_MAX ? INT_MAX : (int)r;
      };
#pragma unroll
      for (int qb = 0; qb < NQB; ++qb) {
        s0[qb] = (MODE == MODE_STATS && p.loss == KGEB_LOSS_KL) ? -INFINITY : 0.f;
        s1[qb] = s2[qb] = 0.f;
        tsc[qb] = 0.f; trel[qb] = -1; fcur[qb] = fend[qb] = ucur[qb] = uend[qb] = 0; fnxt[qb] = unxt[qb] = INT_MAX;
#pragma unroll
        for (int k = 0; k < 6; ++k) cnt[qb][k] = 0;
        if (MODE == MODE_RANK) {
          const int64_t r = (qg * NQB + qb) * TILE + trow;
          if (qb < nqb && r < p.B) {
            float t = p.true_score[r];
            tsc[qb] = (t != t) ? -INFINITY : t;
            const int64_t te = load_index(p.true_ent, p.idx64, r) - chunk_ent0;
            trel[qb] = (te < 0 || te >= (int64_t)INT_MAX) ? -1 : (int)te;
            if (p.f_off) {
              fend[qb] = (int)p.f_off[r + 1];
              fcur[qb] = (int)lb_i64(p.f_col, p.f_off[r], p.f_off[r + 1], chunk_ent0);
              fnxt[qb] = rel_of(p.f_col, fcur[qb], fend[qb]);
            }
            if (p.t_off) {
              uend[qb] = (int)p.t_off[r + 1];
              ucur[qb] = (int)lb_i64(p.t_col, p.t_off[r], p.t_off[r + 1], chunk_ent0);
              unxt[qb] = rel_of(p.t_col, ucur[qb], uend[qb]);
            }
          }
        }
      }

      for (int64_t t = t0; t < t1; ++t) {
        mbar_wait(&t_full[buf], tphase[buf]);
        tphase[buf] ^= 1;
        tc_fence_after();
        const int ncols = (int)min((int64_t)TILE, p.n_ent - t * TILE);  // valid B-tile rows (columns of S)
        const int nv = min(COLS_PER_WARP, ncols - cw0);                  // valid columns of this warp (may be <= 0)
#pragma unroll
        for (int qb = 0; qb < NQB; ++qb) {
          if (qb >= nqb) continue;
          const uint32_t acc = lane_addr + (uint32_t)((buf * NQB + qb) * TILE);
          const int64_t qrow0 = (qg * NQB + qb) * TILE;
          if (MODE == MODE_SCORES) {
            // lanes = entities of tile t, columns = query rows of block qb
            const int64_t e = t * TILE + trow;
            if (qrow0 + cw0 < p.B) {  // warp-uniform
              float v[32];
              tmem_ld32(acc + cw0, v);
              if (e < p.n_ent) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  const int64_t q = qrow0 + cw0 + c;
                  if (q < p.B) p.out[q * p.ld + p.col_off + e] = v[c];
                }
              }
            }
          } else if (MODE == MODE_STATS) {
            if (nv > 0) {
              constexpr float kLog2e = 1.4426950408889634f;
              float v[32];
              tmem_ld32(acc + cw0, v);
              // sum of the raw scores first: columns beyond the table were zero-filled by TMA (score exactly 0), so no
              // masking is needed; afterwards the last tile (warp-uniform test) turns them into -inf, the neutral
              // element of everything below
              float a2 = 0.f;
#pragma unroll
              for (int c = 0; c < 32; ++c) a2 += v[c];
              if (nv < 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c >= nv) v[c] = -INFINITY;
              }
              if (p.loss == KGEB_LOSS_KL) {
                // online max / sum-exp, 5 instructions per element: FADD | FMNMX | FFMA, EX2, FADD
                float mx = s0[qb];
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, v[c]);
                const float mxl = mx * kLog2e;
                float l = (s0[qb] == -INFINITY) ? 0.f : s1[qb] * ex2_ftz(fmaf(s0[qb], kLog2e, -mxl));
#pragma unroll
                for (int c = 0; c < 32; ++c) l += ex2_ftz(fmaf(v[c], kLog2e, -mxl));   // ex2(-inf) = 0
                s0[qb] = mx;
                s1[qb] = l;
                s2[qb] += a2;
              } else {
                // softplus(z) = max(z,0) + ln2 * lg2(1 + 2^(-|z| log2e)); padded columns (z = -inf) contribute 0
                // (KGEB_LG2_GROUP > 1: one lg2 per group of columns, of the product of their 1 + e^-|z| factors in [1, 2])
                float a_lg = 0.f, a_mx = 0.f, prod = 1.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  const float z = v[c] + p.offset;
                  const float a = 1.f + ex2_ftz(fabsf(z) * -kLog2e);
                  if (KGEB_LG2_GROUP > 1) {
                    prod *= a;
                    if (((c + 1) % KGEB_LG2_GROUP) == 0) {
                      a_lg += lg2_ftz(prod);
                      prod = 1.f;
                    }
                  } else {
                    a_lg += lg2_ftz(a);
                  }
                  a_mx += fmaxf(z, 0.f);
                }
                s0[qb] += fmaf(0.69314718f, a_lg, a_mx);
                s2[qb] += fmaf((float)nv, p.offset, a2);
              }
            }
          } else {  // MODE_RANK
            const float ts = tsc[qb];
            const int rel0 = (int)((t - t0) * TILE);        // first column of this tile, relative to the chunk
            if (nv > 0) {
              float v[32];
              tmem_ld32(acc + cw0, v);
              // fast pass (3 instructions per element): rank count and an "anything unusual" flag
              // x > ts  <=>  sign bit of (ts - x), added with one LEA.HI; `plain` stays true while every x is
              // ordered and different from ts (a NaN or an infinite ts - x = inf - inf also sends the row to the
              // exact pass)
              unsigned g0 = 0, g1 = 0;
              bool plain = true;
#pragma unroll
              for (int c = 0; c < 32; c += 2) {
                g0 += __float_as_uint(ts - v[c]) >> 31;
                g1 += __float_as_uint(ts - v[c + 1]) >> 31;
                plain = plain & ((v[c] < ts) | (v[c] > ts)) & ((v[c + 1] < ts) | (v[c + 1] > ts));
              }
              int gt = (int)(g0 + g1);
              const bool eq = !plain;
              // exact pass for the rare rows that need it: a tie, a -inf true score (NaN / filtered candidates tie
              // with it), the true entity's own column (overwritten with the true score, entity_ranking.py:170-177),
              // or the ragged last tile
              const bool own = (unsigned)(trel[qb] - (rel0 + cw0)) < 32u;
              int ties = 0;
              if (eq || own || nv < 32 || ts == -INFINITY) {
                gt = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c < nv) {
                    float x = v[c];
                    if (rel0 + cw0 + c == trel[qb]) x = ts;
                    if (x != x) x = -INFINITY;
                    gt += (x > ts);
                    ties += (x == ts);
                  }
              }
              cnt[qb][0] += gt;
              cnt[qb][1] += ties;
            }
            // corrections for filtered candidates (their score becomes -inf): warp-cooperative, one filter entry per
            // iteration; the value is re-read from the same TMEM accumulator so comparisons are bit-consistent with
            // the raw pass.  Every column part walks the row's entries of this tile and handles those in its columns.
            const int rel_end = rel0 + ncols;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              const int64_t* col = which ? p.t_col : p.f_col;
              int& cur = which ? ucur[qb] : fcur[qb];
              int& nxt = which ? unxt[qb] : fnxt[qb];
              const int end = which ? uend[qb] : fend[qb];
              if (!__any_sync(0xffffffffu, nxt < rel_end)) continue;
              int prev = -1;
              while (true) {
                int loc = -1;
                while (nxt < rel_end) {
                  const int l = nxt - rel0;
                  if ((l / COLS_PER_WARP) == part && nxt != prev && nxt != trel[qb]) { loc = l; break; }
                  prev = nxt;
                  ++cur;
                  nxt = rel_of(col, cur, end);
                }
                const unsigned mask = __ballot_sync(0xffffffffu, loc >= 0);
                if (mask == 0) break;
                const int src = __ffs(mask) - 1;
                const int sloc = __shfl_sync(0xffffffffu, loc, src);
                float x = tmem_ld1(acc + (uint32_t)sloc);
                if (lane == src) {
                  prev = nxt;
                  ++cur;
                  nxt = rel_of(col, cur, end);
                  if (x != x) x = -INFINITY;
                  cnt[qb][2 + 2 * which] -= (x > ts);
                  cnt[qb][3 + 2 * which] += (ts == -INFINITY) - (x == ts);
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&t_empty[buf]);
        buf ^= 1;
      }

      // job results
      if (MODE == MODE_STATS) {
        // exchange the per-part statistics through shared memory; part 0 combines them in a fixed order
#pragma unroll
        for (int qb = 0; qb < NQB; ++qb) {
          float* c = comb + (((size_t)part * 2 + qb) * TILE + trow) * 3;
          c[0] = s0[qb]; c[1] = s1[qb]; c[2] = s2[qb];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
        if (part == 0) {
#pragma unroll
          for (int qb = 0; qb < NQB; ++qb) {
            const int64_t r = (qg * NQB + qb) * TILE + trow;
            if (qb >= nqb || r >= p.B) continue;
            float a0 = s0[qb], a1 = s1[qb], a2 = s2[qb];
            for (int pp = 1; pp < EPQ; ++pp) {
              const float* c = comb + (((size_t)pp * 2 + qb) * TILE + trow) * 3;
              if (p.loss == KGEB_LOSS_KL) {
                const float mx = fmaxf(a0, c[0]);
                const float x = (a0 == -INFINITY) ? 0.f : a1 * __expf(a0 - mx);
                const float y = (c[0] == -INFINITY) ? 0.f : c[1] * __expf(c[0] - mx);
                a0 = mx;
                a1 = x + y;
              } else {
                a0 += c[0];
              }
              a2 += c[2];
            }
            float* o = p.partial + (ch * p.B + r) * 4;
            o[0] = a0; o[1] = a1; o[2] = a2; o[3] = 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");  // comb[] may be rewritten by the next job
      } else if (MODE == MODE_RANK) {
#pragma unroll
        for (int qb = 0; qb < NQB; ++qb) {
          const int64_t r = (qg * NQB + qb) * TILE + trow;
          if (qb >= nqb || r >= p.B) continue;
          unsigned long long* c = p.counts + r * 6;
          const long long rr = cnt[qb][0], rt = cnt[qb][1];
          atomicAdd(c + 0, (unsigned long long)rr);
          atomicAdd(c + 1, (unsigned long long)rt);
          atomicAdd(c + 2, (unsigned long long)(rr + cnt[qb][2]));
          atomicAdd(c + 3, (unsigned long long)(rt + cnt[qb][3]));
          atomicAdd(c + 4, (unsigned long long)(rr + cnt[qb][4]));
          atomicAdd(c + 5, (unsigned long long)(rt + cnt[qb][5]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int d, int box_rows, bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return KGEB_ERR_CUDA;
  }
  const int esize = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for [%lld x %d] at %p", (int)r, (long long)rows, d, base);
    return KGEB_ERR_CUDA;
  }
  return KGEB_OK;
}

static int check_operands(const void* Q, const void* table, int d, bool bf16 = false) {
  KGEB_REQUIRE(d % (bf16 ? 8 : 4) == 0, "tensor tiles need an embedding dim that is a multiple of %d (got %d)",
               bf16 ? 8 : 4, d);
  KGEB_REQUIRE(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(table)) & 15) == 0,
               "TF32 tiles need 16-byte aligned operands");
  if (d > 256) {
    set_error("TF32 tiles are built for entity dim <= 256 (got %d)", d);
    return KGEB_ERR_UNSUPPORTED;
  }
  return KGEB_OK;
}

struct Plan {
  bool bf16;
  int nqb;  // query blocks resident per job (1 or 2)
  Params p;
  size_t smem;
};

static Plan make_plan(int64_t B, int d, int64_t n_ent, int64_t e_lo, bool bf16 = false) {
  Plan pl;
  const int slab_k = bf16 ? 64 : 32;
  const int ks = (d + slab_k - 1) / slab_k;
  pl.bf16 = bf16;
  const int64_t n_qblocks = (B + TILE - 1) / TILE;
  pl.nqb = (n_qblocks >= 2 && ks <= 4) ? 2 : 1;
  const size_t q_bytes = (size_t)pl.nqb * ks * SLAB_BYTES;
  const size_t fixed = 1024 /*align*/ + 512 /*barriers*/ + COMB_BYTES;
  int stages = (int)((SMEM_BUDGET - fixed - q_bytes) / SLAB_BYTES);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  Params& p = pl.p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.n_ent = n_ent; p.e_lo = e_lo; p.d = d; p.ks = ks; p.stages = stages;
  p.n_qgroups = (n_qblocks + pl.nqb - 1) / pl.nqb;
  p.n_tiles = (n_ent + TILE - 1) / TILE;
  int64_t chunks = kNumSMs / p.n_qgroups;
  if (chunks < 1) chunks = 1;
  if (chunks > p.n_tiles) chunks = p.n_tiles;
  p.tiles_per_chunk = (p.n_tiles + chunks - 1) / chunks;
  p.chunks = (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  pl.smem = q_bytes + (size_t)stages * SLAB_BYTES + fixed;
  return pl;
}

template <int MODE>
static int launch(const Plan& pl, const CUtensorMap& mq, const CUtensorMap& mw, cudaStream_t st) {
  const int64_t jobs = pl.p.n_qgroups * pl.p.chunks;
  const int grid = (int)(jobs < kNumSMs ? jobs : kNumSMs);
  cudaError_t e;
#define KGEB_TC_LAUNCH(NQB_, BF_)                                                                                     \
  {                                                                                                                   \
    e = cudaFuncSetAttribute(tc_tiles_kernel<MODE, NQB_, BF_>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                             (int)pl.smem);                                                                           \
    if (e != cudaSuccess) return cuda_status(e, "tc smem attribute");                                                 \
    tc_tiles_kernel<MODE, NQB_, BF_><<<grid, NUM_THREADS, pl.smem, st>>>(mq, mw, pl.p);                               \
  }
  if (pl.nqb == 2) {
    if (pl.bf16) KGEB_TC_LAUNCH(2, true) else KGEB_TC_LAUNCH(2, false)
  } else {
    if (pl.bf16) KGEB_TC_LAUNCH(1, true) else KGEB_TC_LAUNCH(1, false)
  }
#undef KGEB_TC_LAUNCH
  KGEB_LAUNCH_CHECK("tc_tiles_kernel");
  return KGEB_OK;
}

__global__ void stats_finalize_kernel(int loss, const float* __restrict__ partial, int64_t B, int64_t chunks,
                                      const float* __restrict__ label_dot, float* __restrict__ rowstat) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= B) return;
  float s0 = (loss == KGEB_LOSS_KL) ? -INFINITY : 0.f, s1 = 0.f, s2 = 0.f;
  for (int64_t c = 0; c < chunks; ++c) {
    const float* q = partial + (c * B + r) * 4;
    if (loss == KGEB_LOSS_KL) {
      float mx = fmaxf(s0, q[0]);
      float a = (s0 == -INFINITY) ? 0.f : s1 * __expf(s0 - mx);
      float b = (q[0] == -INFINITY) ? 0.f : q[1] * __expf(q[0] - mx);
      s0 = mx;
      s1 = a + b;
    } else {
      s0 += q[0];
    }
    s2 += q[2];
  }
  float* o = rowstat + r * 4;
  o[0] = s0; o[1] = s1; o[2] = s2; o[3] = label_dot[r];
}

// exact-fp32 x_ij at the label entries of this shard, one warp per entry (+ add_per_entry), then ordered row sums
__global__ void __launch_bounds__(256)
label_entry_dot_kernel(const float* __restrict__ Q, int64_t B, int d, const float* __restrict__ table, int64_t e_lo,
                       int64_t n_ent, const int64_t* __restrict__ lab_off, const int64_t* __restrict__ lab_col,
                       float add_per_entry, int64_t nnz, float* __restrict__ entry_dot) {
  const int lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nnz) return;
  int64_t q = -1;
  if (lane == 0 && i < lab_off[B]) {
    int64_t lo = 0, hi = B;
    while (lo < hi) {
      int64_t mid = (lo + hi + 1) >> 1;
      if (lab_off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    q = lo;
  }
  q = __shfl_sync(0xffffffffu, q, 0);
  float out = 0.f;
  if (q >= 0) {
    const int64_t c = lab_col[i] - e_lo;
    if (c >= 0 && c < n_ent) {
      const float* qr = Q + q * d;
      const float* er = table + c * d;
      float acc = 0.f;
      for (int k = lane; k < d; k += 32) acc = fmaf(qr[k], __ldg(er + k), acc);
      out = warp_sum(acc) + add_per_entry;
    }
  }
  if (lane == 0) entry_dot[i] = out;
}
__global__ void __launch_bounds__(256)
label_row_sum_kernel(const float* __restrict__ entry_dot, const int64_t* __restrict__ lab_off, int64_t B,
                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  float acc = 0.f;
  if (lab_off)
    for (int64_t i = lab_off[r] + lane; i < lab_off[r + 1]; i += 32) acc += entry_dot[i];
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc;
}

}  // namespace tc

int tc_score_all(const float* Q, int64_t B, int d, const float* table, int64_t m, float* out, int64_t ld,
                 int64_t col_off, cudaStream_t st) {
  int rc = tc::check_operands(Q, table, d);
  if (rc) return rc;
  tc::Plan pl = tc::make_plan(B, d, m, 0);
  pl.p.out = out; pl.p.ld = ld; pl.p.col_off = col_off;
  CUtensorMap mq, mw;
  if ((rc = tc::make_map(&mq, Q, B, d, tc::TILE)) || (rc = tc::make_map(&mw, table, m, d, tc::TILE))) return rc;
  return tc::launch<tc::MODE_SCORES>(pl, mq, mw, st);
}

int64_t tc_stats_partial_bytes(int64_t B, int64_t nnz) {
  return (int64_t)kNumSMs * B * 4 * (int64_t)sizeof(float) + B * 4 + (nnz < 1 ? 1 : nnz) * 4 + 512;
}

// math: KGEB_MATH_TF32 reads the fp32 Q / table in place; KGEB_MATH_BF16 reads the bf16 mirrors Qb / tableb
int tc_fused_fwd(int loss, int math, const float* Q, const void* Qb, int64_t B, int d, const float* table,
                 const void* tableb, int64_t e_lo, int64_t n_ent, const int64_t* lab_off, const int64_t* lab_col,
                 int64_t nnz, float ls_keep, float ls_add, float offset, float* rowstat, void* ws, int64_t ws_bytes,
                 cudaStream_t st) {
  const bool bf16 = (math == KGEB_MATH_BF16);
  int rc = tc::check_operands(bf16 ? Qb : (const void*)Q, bf16 ? tableb : (const void*)table, d, bf16);
  if (rc) return rc;
  KGEB_REQUIRE(ws_bytes >= tc_stats_partial_bytes(B, nnz), "fused_fwd(tensor tiles): workspace too small");
  float* partial = reinterpret_cast<float*>(ws);
  tc::Plan pl = tc::make_plan(B, d, n_ent, e_lo, bf16);
  float* label_dot = partial + (int64_t)kNumSMs * B * 4;
  float* entry_dot = label_dot + B;
  pl.p.loss = loss; pl.p.offset = offset; pl.p.partial = partial;
  if (n_ent > 0) {
    CUtensorMap mq, mw;
    if ((rc = tc::make_map(&mq, bf16 ? Qb : (const void*)Q, B, d, tc::TILE, bf16)) ||
        (rc = tc::make_map(&mw, bf16 ? tableb : (const void*)table, n_ent, d, tc::TILE, bf16)))
      return rc;
    if ((rc = tc::launch<tc::MODE_STATS>(pl, mq, mw, st))) return rc;
  }
  if (nnz > 0 && lab_off) {
    tc::label_entry_dot_kernel<<<(unsigned)((nnz + 7) / 8), 256, 0, st>>>(
        Q, B, d, table, e_lo, n_ent, lab_off, lab_col, loss == KGEB_LOSS_BCE ? offset : 0.f, nnz, entry_dot);
    KGEB_LAUNCH_CHECK("label_entry_dot");
  }
  tc::label_row_sum_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(entry_dot, nnz > 0 ? lab_off : nullptr, B, label_dot);
  KGEB_LAUNCH_CHECK("label_row_sum");
  tc::stats_finalize_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(loss, partial, B, n_ent > 0 ? pl.p.chunks : 0,
                                                                         label_dot, rowstat);
  KGEB_LAUNCH_CHECK("stats_finalize");
  return KGEB_OK;
}

int tc_rank_count(const float* Q, int64_t nq, int d, const float* table, int64_t e_lo, int64_t n_ent,
                  const float* true_score, const void* true_ent, int idx64, const int64_t* f_off,
                  const int64_t* f_col, const int64_t* t_off, const int64_t* t_col, int64_t* counts,
                  cudaStream_t st) {
  int rc = tc::check_operands(Q, table, d);
  if (rc) return rc;
  tc::Plan pl = tc::make_plan(nq, d, n_ent, e_lo);
  pl.p.true_score = true_score; pl.p.true_ent = true_ent; pl.p.idx64 = idx64;
  pl.p.f_off = f_off; pl.p.f_col = f_col; pl.p.t_off = t_off; pl.p.t_col = t_col;
  pl.p.counts = reinterpret_cast<unsigned long long*>(counts);
  CUtensorMap mq, mw;
  if ((rc = tc::make_map(&mq, Q, nq, d, tc::TILE)) || (rc = tc::make_map(&mw, table, n_ent, d, tc::TILE))) return rc;
  return tc::launch<tc::MODE_RANK>(pl, mq, mw, st);
}

}  // namespace kgeb
