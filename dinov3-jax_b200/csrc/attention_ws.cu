// Warp-specialised, persistent multi-head self-attention forward for the short DINOv3 sequences (N <= 256 tokens per
// crop, head_dim 64): replaces flax `nn.dot_product_attention(q, k, v)` at dinov3_jax/layers/attention.py:116.
//
// One CTA per SM loops over work items (crop group, head).  Roles inside the CTA (320 threads):
//   warp 0      TMA producer: Q tiles, K and V of the NEXT item land in a 2-stage shared-memory ring while the
//               current item is being computed;
//   warp 1      tcgen05 issuer: S = Q K^T (SS form) into one of two tensor-memory slots, and O = P V with P read
//               straight from tensor memory (TS form: no shared-memory round trip for the probabilities);
//   warps 2-5 / 6-9   two softmax warpgroups, one per slot: each thread owns one query row (tcgen05.ld 32x32b), takes the
//               row maximum, writes P = exp2(..) as packed bf16 back over the consumed S columns (tcgen05.st), and
//               after the PV product normalises and stores its output row and the log-sum-exp.
// The two slots ping-pong: the tensor core works on one query tile while the other tile is in its exponentials, which
// are the bound of this kernel (4 * N^2 * 64 flop vs N^2 ex2 per (crop, head): 16 MUFU lanes per SM).
#include "ptx.cuh"
#include <cstdlib>
#include "d3_internal.h"

namespace d3 {

constexpr float LOG2E_WS = 1.4426950408889634f;

struct AttnWsShape {
  int G, span, n_crops, N, Nkp, nbox, box_rows, D, H;
  float scale;
  int nQ;           // query tiles per item
  int n_groups;     // crop groups (items = n_groups * H)
  int slot_w;       // tensor-memory columns per slot (128 or 256)
  int n_slot;       // 512 / slot_w: query tiles in flight (2 or 4)
  int o_off;        // column of the O accumulator inside a slot
  int mid;          // key column where the upper thread of a row takes over (multiple of 16); its packed P starts at column mid
  int stage_bytes;  // shared memory per pipeline stage
  int n_stage;      // shared-memory stages (items prefetched): 2 .. 4
  int nq_sh, ns_sh, nst_sh;   // log2 of nQ, n_slot, n_stage (all powers of two: index math by shifts, no division)
  unsigned div_magic;  // floor(q / N) == (q * div_magic) >> 16 for every q < 256 (checked on the host)
  int dbg;          // D3_ATTN_DEBUG bits (timing experiments only, results wrong): 1 skip max pass, 2 skip exponentials,
                    // 4 skip the O read-out / stores
};

// optional clock64() trace of CTA 0 (tools/attn_ws_trace.py): [unit < 16][event < 20]
__device__ long long* g_ws_trace = nullptr;
__device__ __forceinline__ void ws_mark(int u, int ev) {
  if (g_ws_trace && blockIdx.x == 0 && u < 16) g_ws_trace[u * 20 + ev] = clock64();
}
void attn_ws_set_trace(long long* buf) { cudaMemcpyToSymbol(g_ws_trace, &buf, sizeof(buf)); }

// maximum of 16 fp32 values held as raw bits: a 4-level tree (the serial form is a 16-deep dependency chain)
__device__ __forceinline__ float max16(const uint32_t (&v)[16]) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 8]));
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = fmaxf(a[i], a[i + 4]);
  return fmaxf(fmaxf(a[0], a[2]), fmaxf(a[1], a[3]));
}

constexpr int WS_MAX_STAGE = 4, WS_MAX_SLOT = 4;
constexpr int WS_THREADS = 608;   // warp 0: TMA, warps 1-2: tensor-core issuers (even / odd units), warps 3-18: softmax

// Units (query tiles) are numbered u = item_local * nQ + qt in the order a CTA meets them.  Unit u lives in slot
// u % n_slot; even units are issued by warp 1 and handled by warpgroup 0, odd units by warp 2 / warpgroup 1, so the two
// chains  S -> row math -> PV -> read-out  never wait on each other's program order.
__global__ void __launch_bounds__(WS_THREADS, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   __nv_bfloat16* __restrict__ O, float* __restrict__ LSE, const AttnWsShape sh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = sh.Nkp * 128;
  // stage layout: [Q tiles nQ x 16 KB][K Nkp x 128 B][V Nkp x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.n_stage * sh.stage_bytes);
  uint64_t* bar_full = bars;                       // [stage] TMA bytes of a stage have landed
  uint64_t* bar_empty = bars + WS_MAX_STAGE;       // [stage] every MMA that reads the stage has retired (5 arrivals per unit: PV commit + 4 read-out warps)
  uint64_t* bar_s = bar_empty + WS_MAX_STAGE;      // [slot] S is in tensor memory
  uint64_t* bar_p = bar_s + WS_MAX_SLOT;           // [slot] P has been written (4 warp arrivals)
  uint64_t* bar_o = bar_p + WS_MAX_SLOT;           // [slot] O is complete
  uint64_t* bar_free = bar_o + WS_MAX_SLOT;        // [slot] O has been read out: the slot may be overwritten (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_free + WS_MAX_SLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = sh.n_groups * sh.H;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items b, b+grid, ...
  const int n_units = my_items * sh.nQ;
  const int NS = sh.n_slot, NST = sh.n_stage;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < WS_MAX_STAGE; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], sh.nQ * 9);          // per unit: the PV commit + the eight read-out warps
    }
    for (int i = 0; i < WS_MAX_SLOT; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 8);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_free[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int j = 0; j < my_items; ++j) {
        const int item = blockIdx.x + j * gridDim.x;
        const int c = item / sh.H, h = item - c * sh.H;
        const int st = (j & (NST - 1));
        if (j >= NST) mbar_wait(&bar_empty[st], ((j >> sh.nst_sh) - 1) & 1);
        uint8_t* base = smem + st * sh.stage_bytes;
        const int row_base = c * sh.span;
        mbar_expect_tx(&bar_full[st], sh.nQ * 16384 + 2 * kv_bytes);
        for (int qt = 0; qt < sh.nQ; ++qt) tma_load_2d(&tmQ, &bar_full[st], base + qt * 16384, h * 64, row_base + qt * 128);
        uint8_t* sK = base + sh.nQ * 16384;
        uint8_t* sV = sK + kv_bytes;
        for (int b = 0; b < sh.nbox; ++b) {
          tma_load_2d(&tmKV, &bar_full[st], sK + b * sh.box_rows * 128, sh.D + h * 64, row_base + b * sh.box_rows);
          tma_load_2d(&tmKV, &bar_full[st], sV + b * sh.box_rows * 128, 2 * sh.D + h * 64, row_base + b * sh.box_rows);
        }
      }
    }
    __syncwarp();
  } else if (warp <= 2) {
    // ------------------------------------------------------------------------------------------ tensor-core issuers
    if (elect_one()) {
      const int par = warp - 1;                       // this warp issues the units u with u % 2 == par
      const uint32_t idesc_s = umma_idesc_bf16(128, sh.Nkp, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      const int ksteps = sh.Nkp / 16;
      const int lag = NS > 2 ? 2 : 0;                 // with 4 slots the PV of unit u-2 is issued after S of unit u
      auto issue_s = [&](int u) {
        const int j = u >> sh.nq_sh, qt = u - (j << sh.nq_sh), st = (j & (NST - 1)), slot = (u & (NS - 1));
        mbar_wait(&bar_full[st], (j >> sh.nst_sh) & 1);
        if (u >= NS) mbar_wait(&bar_free[slot], ((u >> sh.ns_sh) - 1) & 1);
        tc_fence_after();
        ws_mark(u, 0);
        const uint8_t* base = smem + st * sh.stage_bytes;
        const uint64_t qd = umma_desc_sw128(smem_u32(base + qt * 16384), 16, 1024);
        const uint64_t kd = umma_desc_sw128(smem_u32(base + sh.nQ * 16384), 16, 1024);
        const uint32_t tS = tmem + slot * sh.slot_w;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bar_s[slot]);
        ws_mark(u, 1);
      };
      auto issue_pv = [&](int v) {
        const int j = v >> sh.nq_sh, st = (j & (NST - 1)), slot = (v & (NS - 1));
        mbar_wait(&bar_p[slot], (v >> sh.ns_sh) & 1);
        tc_fence_after();
        ws_mark(v, 2);
        const uint8_t* base = smem + st * sh.stage_bytes;
        const uint64_t vd = umma_desc_sw128(smem_u32(base + sh.nQ * 16384 + kv_bytes), 8192, 1024);
        const uint32_t tP = tmem + slot * sh.slot_w;
        const uint32_t tO = tP + sh.o_off;
        // packed P: keys below sh.mid from column 0, keys from sh.mid on from column sh.mid (8 columns per 16 keys)
        const int hi_shift = sh.mid - (sh.mid >> 1);
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ts(tO, tP + k * 8 + (16 * k >= sh.mid ? hi_shift : 0), vd + (uint64_t)(k * 128), idesc_pv, k > 0 ? 1u : 0u);
        umma_commit(&bar_o[slot]);
        umma_commit(&bar_empty[st]);                  // nQ arrivals (one per unit of the item) release the stage
        ws_mark(v, 3);
      };
      int u = par;
      for (; u < n_units; u += 2) {
        issue_s(u);
        if (u - lag >= 0) issue_pv(u - lag);
      }
      if (lag && u - lag < n_units && u - lag >= 0) issue_pv(u - lag);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ softmax warpgroups
    // 256 threads per warpgroup: two threads per query row.  Warp w of the CTA reaches tensor-memory lanes
    // 32*(w%4)..+31, so the two warps with equal w%4 inside a warpgroup share a row quadrant and split its key columns
    // (ch = 0 / 1).  Row maximum and row sum are exchanged through shared memory + a 256-thread named barrier.
    const int wg = (warp - 3) >> 3;                      // parity of the units this warpgroup serves
    const int ch = ((warp - 3) >> 2) & 1;                // column half
    const int r = (warp & 3) * 32 + lane;                // tile row == tensor-memory lane
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    const float cs = sh.scale * LOG2E_WS;
    const int lag = NS > 2 ? 2 : 0;
    float* red_mx = reinterpret_cast<float*>(tmem_slot + 4) + wg * 256;     // [2 halves][128 rows], per warpgroup
    float* red_sum = reinterpret_cast<float*>(tmem_slot + 4) + 512;         // [slot][2 halves][128 rows]
    // state of the unit whose read-out is still pending (4-slot mode defers it behind the next unit's row math)
    float p_mx = 0.f;
    int p_u = -1;
    const bool tracer = ((warp - 3) & 7) == 1 && lane == 0;      // the ch = 0 warp that owns tile rows 0..31

    auto wg_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); };

    auto readout = [&](int u, float mx) {
      const int j = u >> sh.nq_sh, qt = u - (j << sh.nq_sh);
      const int item = blockIdx.x + j * gridDim.x;
      const int c = item / sh.H, h = item - c * sh.H;
      const int q_abs = qt * 128 + r;
      const int g = min((int)((q_abs * sh.div_magic) >> 16), sh.G - 1);
      const int klo = g * sh.N;
      const bool q_valid = (q_abs < sh.span) && (c * sh.G + g < sh.n_crops);
      const bool warp_live = qt * 128 + (warp & 3) * 32 < sh.span;
      const int slot = (u & (NS - 1));
      const uint32_t tS = tmem + slot * sh.slot_w + t_lane;
      mbar_wait(&bar_o[slot], (u >> sh.ns_sh) & 1);
      tc_fence_after();
      if (tracer) ws_mark(u, 7);
      const bool active = warp_live && !(sh.dbg & 4);
      const float sum = red_sum[slot * 256 + r] + red_sum[slot * 256 + 128 + r];   // written before bar_p was signalled
      const float inv = 1.f / sum;
      // this thread's 32 of the row's 64 output columns, 16 at a time straight into the staging tile: the rows go out
      // through the query tile of this unit in shared memory (dead since S = Q K^T completed); each lane parks its
      // 64 bytes (16-byte chunks XOR-swizzled by row), then every store instruction of the warp covers eight half-rows
      // of 64 contiguous bytes instead of 32 row-strided 16-byte pieces.
      uint8_t* stg = smem + ((j & (NST - 1))) * sh.stage_bytes + qt * 16384 + (warp & 3) * 4096;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[16];
        if (active) {
          tmem_ld16(tS + sh.o_off + ch * 32 + half * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 2; ++i)
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((ch * 4 + half * 2 + i) ^ (lane & 7)) << 4)) = make_uint4(
                pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_free[slot]);     // the slot may take its next S while the rows go out to memory
      if (tracer) ws_mark(u, 8);
      __syncwarp();
      if (active) {
        const int chunk = ch * 4 + (lane & 3);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + (lane >> 2);
          const int q = qt * 128 + (warp & 3) * 32 + row;
          const int gq = min((int)((q * sh.div_magic) >> 16), sh.G - 1);
          if (q < sh.span && c * sh.G + gq < sh.n_crops) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(O + (size_t)(c * sh.span + q) * sh.D + h * 64 + chunk * 8) = v4;
          }
        }
        if (q_valid && LSE && ch == 0) LSE[((size_t)(c * sh.G + g) * sh.H + h) * sh.N + (q_abs - klo)] = mx * sh.scale + logf(sum);
      }
      // write-after-read on the staging rows: the shared-memory reads above have returned their data; the arrive /
      // wait pair on bar_empty orders them before the TMA load that refills this stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[(j & (NST - 1))]);
      if (tracer) ws_mark(u, 9);
    };

    for (int u = wg; u < n_units; u += 2) {
      const int j = u >> sh.nq_sh, qt = u - (j << sh.nq_sh);
      const int q0 = qt * 128;
      const int q_abs = q0 + r;
      const int g = min((int)((q_abs * sh.div_magic) >> 16), sh.G - 1);
      const int klo = g * sh.N, khi = klo + sh.N;
      const bool warp_live = q0 + (warp & 3) * 32 < sh.span;
      // 16-column chunks any row of this warp needs (block-diagonal packing of short crops: a warp spans <= 2 crops);
      // the lower thread of a row takes the chunks below sh.mid, the upper thread those from sh.mid on
      const int wlo = __reduce_min_sync(0xffffffffu, klo), whi = __reduce_max_sync(0xffffffffu, khi);
      const int r_beg = (wlo >> 4) << 4, r_end = min(sh.Nkp, (whi + 15) & ~15);
      const int h_lo = ch ? sh.mid : 0, h_hi = ch ? sh.Nkp : sh.mid;       // this thread's half of the key columns
      const int c_beg = max(r_beg, h_lo), c_end = min(r_end, h_hi);          // may be empty (c_beg >= c_end)
      // a 16-column chunk is "full" when it lies inside the key range of EVERY row of the warp: a warp-uniform test, so
      // the unmasked fast path is a real branch (a per-lane condition gets if-converted: both variants execute)
      const int f_lo = __reduce_max_sync(0xffffffffu, klo), f_hi = __reduce_min_sync(0xffffffffu, khi);
      const int slot = (u & (NS - 1));
      const uint32_t tS = tmem + slot * sh.slot_w + t_lane;
      // packed P of key column c lives at column p_col(c): the lower half in place from column 0, the upper half in
      // place from column sh.mid (so each thread's writes trail its own reads and never touch the other half's S)
      const int p_base = ch ? sh.mid - (sh.mid >> 1) : 0;                    // p_col(c0) = p_base + c0 / 2
      float* rsum = red_sum + slot * 256;

      mbar_wait(&bar_s[slot], (u >> sh.ns_sh) & 1);
      // The two chains are symmetric, so left alone they run in lockstep (both in their exponentials, then both waiting
      // on the tensor core).  Holding the second warpgroup back once, by about half an item period, keeps one chain in
      // its MUFU-bound phase while the other is in its tensor-core / read-out phases (measured 9.3k -> 7.8k cycles per
      // item at N = 197); nothing re-synchronises them afterwards.
      if (u == 1 && sh.nQ == 2 && !(sh.dbg & 128)) __nanosleep(2800);
      tc_fence_after();
      if (tracer) ws_mark(u, 4);
      float mx = -3.0e38f, sum = 0.f;
      uint32_t va[16], vb[16];
      // ---- pass 1: maximum of this thread's half of the row
      if (warp_live && !(sh.dbg & 1) && c_beg < c_end) {
        auto max_chunk = [&](const uint32_t (&v)[16], int c0) {
          if (c0 >= f_lo && c0 + 16 <= f_hi) {
            mx = fmaxf(mx, max16(v));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i >= klo && c0 + i < khi) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        };
        tmem_ld16(tS + c_beg, va);
        tmem_ld_wait();
#pragma unroll 1
        for (int c0 = c_beg; c0 < c_end; c0 += 32) {
          const bool hasB = c0 + 16 < c_end;
          if (hasB) tmem_ld16(tS + c0 + 16, vb);
          max_chunk(va, c0);
          tmem_ld_wait();
          if (hasB) {
            if (c0 + 32 < c_end) tmem_ld16(tS + c0 + 32, va);
            max_chunk(vb, c0 + 16);
            tmem_ld_wait();
          }
        }
      }
      if (sh.dbg & 1) mx = 0.f;
      red_mx[ch * 128 + r] = mx;
      wg_sync();
      mx = fmaxf(red_mx[r], red_mx[128 + r]);
      if (tracer) ws_mark(u, 5);
      const float mxs = mx * cs;
      // ---- pass 2: P = 2^(s*cs - mxs) as packed bf16
      if (warp_live) {
        auto emit = [&](const uint32_t (&v)[16], int c0) {
          uint32_t pw[8];
          if (c0 >= f_lo && c0 + 16 <= f_hi) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), cs, -mxs));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), cs, -mxs));
              sum += p0 + p1;
              pw[i >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = (c0 + i >= klo && c0 + i < khi) ? ex2_approx(fmaf(__uint_as_float(v[i]), cs, -mxs)) : 0.f;
              const float p1 = (c0 + i + 1 >= klo && c0 + i + 1 < khi) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), cs, -mxs)) : 0.f;
              sum += p0 + p1;
              pw[i >> 1] = pack_bf16(p0, p1);
            }
          }
          tmem_st8(tS + p_base + (c0 >> 1), pw);
        };
        const uint32_t zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // chunks of this half outside the warp's key range are zero probabilities; the ones below c_beg are written
        // first (their packed columns lie below every column this thread still has to read)
        for (int c0 = h_lo; c0 < min(max(c_beg, h_lo), h_hi); c0 += 16) tmem_st8(tS + p_base + (c0 >> 1), zeros);
        if (sh.dbg & 2) {
          for (int c0 = c_beg; c0 < c_end; c0 += 16) tmem_st8(tS + p_base + (c0 >> 1), zeros);
          sum = 0.5f;
        } else if (c_beg < c_end) {
          tmem_ld16(tS + c_beg, va);
          tmem_ld_wait();
#pragma unroll 1
          for (int c0 = c_beg; c0 < c_end; c0 += 32) {
            const bool hasB = c0 + 16 < c_end;
            if (hasB) tmem_ld16(tS + c0 + 16, vb);
            tmem_ld_wait();                   // vb is in registers before emit() overwrites packed columns under it
            emit(va, c0);
            if (hasB) {
              if (c0 + 32 < c_end) tmem_ld16(tS + c0 + 32, va);
              tmem_ld_wait();
              emit(vb, c0 + 16);
            }
          }
        }
        for (int c0 = max(c_end, h_lo); c0 < h_hi; c0 += 16) tmem_st8(tS + p_base + (c0 >> 1), zeros);
        tmem_st_wait();
      }
      rsum[ch * 128 + r] = sum;                   // read at read-out time, ordered by the bar_p -> bar_o chain
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[slot]);
      if (tracer) ws_mark(u, 6);

      // ---- O = P V: normalise, store the row and its log-sum-exp (4-slot mode: one unit behind, so that the PV
      // product of this unit runs under the row math of the next one)
      if (lag) {
        if (p_u >= 0) readout(p_u, p_mx);
        p_u = u; p_mx = mx;
      } else {
        readout(u, mx);
      }
    }
    if (lag && p_u >= 0) readout(p_u, p_mx);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

// returns D3_OK and sets *handled = 1 when the shape is served by this kernel
int attn_fwd_ws(const void* qkv, void* o, float* lse, int n_crops, int N, int D, int H, cudaStream_t st, int* handled) {
  *handled = 0;
  if (D != H * 64 || N <= 0 || n_crops <= 0) return D3_OK;
  AttnWsShape s;
  s.N = N; s.D = D; s.H = H; s.scale = 0.125f; s.n_crops = n_crops;
  s.G = (N <= 64) ? (128 / N) : 1;
  if (s.G > n_crops) s.G = n_crops;
  s.span = s.G * N;
  s.nbox = (s.span + 255) / 256;
  const int q = 16 * s.nbox;
  s.Nkp = (s.span + q - 1) / q * q;
  if (s.Nkp > 256) return D3_OK;                       // longer crops: single-pass kernel in attention.cu
  s.box_rows = s.Nkp / s.nbox;
  s.nQ = (s.span + 127) / 128;
  s.n_groups = (n_crops + s.G - 1) / s.G;
  s.slot_w = s.Nkp > 128 ? 256 : 128;
  s.n_slot = 512 / s.slot_w;
  if (s.slot_w == 256) {       // two threads per row split the keys at `mid`; O sits above both packed-P ranges
    s.mid = ((s.Nkp / 16 + 1) / 2) * 16;
    s.o_off = 192;
  } else {                     // narrow slots: the packed P stays contiguous (the second thread of a row only shares the read-out)
    s.mid = s.Nkp;
    s.o_off = 64;
  }
  s.stage_bytes = s.nQ * 16384 + 2 * s.Nkp * 128;
  s.div_magic = (65536u + N - 1) / N;
  for (unsigned qq = 0; qq < 256; ++qq)
    if (((qq * s.div_magic) >> 16) != qq / (unsigned)N) return D3_OK;       // never for N <= 256; falls back if it did
  s.n_stage = (216 * 1024) / s.stage_bytes;
  if (s.n_stage < 2 || s.nQ != 2) return D3_OK;   // one-tile crop groups (96^2 local crops) stay on the single-pass kernel: 48 vs 61 us
  s.n_stage = s.n_stage >= 4 ? 4 : 2;
  s.nq_sh = s.nQ == 2 ? 1 : 0;
  s.ns_sh = s.n_slot == 4 ? 2 : 1;
  s.nst_sh = s.n_stage == 4 ? 2 : 1;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("D3_ATTN_DEBUG"); dbg = e ? atoi(e) : 0; }
    s.dbg = dbg;
  }
  const long T = (long)n_crops * N;
  CUtensorMap tq, tkv;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * D), (cuuint64_t)T};
  cuuint64_t strides[1] = {(cuuint64_t)(3 * D) * 2};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t boxq[2] = {64, 128}, boxkv[2] = {64, (cuuint32_t)s.box_rows};
  int rc;
  if ((rc = encode_tensor_map_2d_bf16(&tq, qkv, dims, strides, boxq, estr))) return rc;
  if ((rc = encode_tensor_map_2d_bf16(&tkv, qkv, dims, strides, boxkv, estr))) return rc;
  const int smem = s.n_stage * s.stage_bytes + 256 + 6144 + 1024;     // barriers + row-reduction scratch + alignment
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    cfg = true;
  }
  const int items = s.n_groups * H;
  const int grid = items < sm_count() ? items : sm_count();
  attn_fwd_ws_kernel<<<grid, WS_THREADS, smem, st>>>(tq, tkv, (__nv_bfloat16*)o, lse, s);
  D3_CHECK_LAUNCH();
  *handled = 1;
  return D3_OK;
}

}  // namespace d3
