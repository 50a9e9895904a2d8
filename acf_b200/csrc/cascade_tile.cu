// cascade_tile.cu -- k_cascade_tile: the depth-2 sliding-window cascade (toolbox/acfDetect1.cpp:84-138) on
// shared-memory channel tiles.
//
// A thread block owns a TILE of Wc x Wr neighbouring windows of one (frame, scale).  The channel footprint of the
// tile -- ((Wc-1) step + modelWd/shrink) columns x ((Wr-1) step + modelHt/shrink) rows x all channels, 80-90 KB --
// is brought into shared memory by ONE 4-D cp.async.bulk.tensor (TMA, dims y | x | channel | frame; out-of-range
// parts of small scales are zero-filled by the copy engine) behind an mbarrier; two blocks share an SM, so one
// block's copy is in flight while the other one decides its windows.  Every feature gather of the cascade is then
// an LDS: ~30 cycles instead of an L2 round trip, and the L2 -> SM traffic per window drops from ~1 KB (three
// 32-byte sectors per tree) to the tile's halo-amplified footprint (~190 B).
//
// Tree records (three tile-local byte offsets, three thresholds, four leaf values: 48 B): trees [0, 64) travel in the
// kernel parameters (constant-bank / uniform-register operands of the unrolled head levels; a shared-memory copy serves
// models with fewer trees), later trees stream through a two-slot shared-memory ring of 64-tree chunks (cp.async.bulk,
// one chunk ahead).  Offsets are tile-local constants because every tile of a launch has
// the same shared-memory pitch, so a gather is base(window) + offset(node): one add, no multiply.
//
// Early-exit compaction is level-synchronous and block-wide: trees are cut into levels [0,4) [4,8) [8,64) [64,128) [128,192)
// ... (ctSegEnd; five head levels [0,4) [4,8) [8,16) [16,32) [32,64) behind ACFB_CASC_HEAD_LEVELS=5); after a level the
// survivors of each 32-window batch are appended (warp ballot, one shared-memory atomic per batch) to the block's survivor
// list, and the next level walks that list 32 entries per warp, so late trees run on (nearly) full warps although most
// windows die after a handful of trees.  A window's score is the reference's sequential float sum h += leaf(t), compared
// with cascThr after every tree, so hit sets, scores and the number of trees evaluated are bit-identical to the CPU path.
// The few windows still alive past tree 64 (hits walk every tree) are handed to k_cascade_tail_win below: one window per
// warp on its own TMA-staged footprint, lanes = trees (k_cascade_tail: the same with global gathers, as a fallback).
#include "kernels.cuh"
#include "tma.cuh"
#include <cuda.h>
#include <algorithm>
#include <cstdio>

namespace acfb
{

#define FULLMASK 0xffffffffu

constexpr int kCtChunk = 64;     // trees per table chunk (resident part and ring slots)
constexpr int kCtRecWords = 12;  // {off0, off1, off2, thr0} {thr1, thr2, leaf0, leaf1} {leaf2, leaf3, -, -}
constexpr int kCtFixedBytes = 3 * kCtChunk * kCtRecWords * 4 + 4 * 8 + 32 * 4; // tables + mbarriers + block scalars

// Level l covers trees [ctSegEnd(l - 1), ctSegEnd(l)).  nHead = 5: [0,4) [4,8) [8,16) [16,32) [32,64), then 64-tree chunks;
// nHead = 3: [0,4) [4,8) [8,64) -- after eight trees few enough windows are left that further compaction saves fewer issue
// slots than its block barriers cost.
__host__ __device__ inline int ctSegEnd(int lvl, int nHead)
{
    if (lvl >= nHead) return kCtChunk * (lvl - nHead + 2);
    return (nHead == 3 && lvl == 2) ? kCtChunk : (4 << lvl);
}

// block scalars in shared memory
enum { kSiTask = 0, kSiExport, kSiFrame, kSiScale, kSiC0, kSiR0, kSiNc, kSiNr, kSiScaleIdx, kSiCnt /* 3 counters */ };

// Run trees [0, nT) of the table at shared address `tab` (48 B per tree: {off0, off1, off2, thr0} {thr1, thr2, leaf0, leaf1}
// {leaf2, leaf3}) on up to 32 windows whose tile-local origin is at shared address `wb`; returns the ballot of the
// survivors (acfDetect1.cpp:100-138: h += leaf; if (h <= cascThr) break).  Three-stage software pipeline, because a step
// is a chain of two dependent shared-memory loads (record -> feature) and the decision: while tree t is decided, the
// features of tree t+1 and the offsets of tree t+2 are in flight.
__device__ __forceinline__ unsigned ctSegment(uint32_t tab, int nT, float cascThr, uint32_t wb, bool valid, float& h, unsigned& nEval)
{
    bool alive = valid;
    const uint32_t tabLast = tab + 48u * (uint32_t)(nT - 1);
    // tree 0: everything; tree 1: offsets
    uint4 P0 = lds128(tab);
    uint4 Q0 = lds128(tab + 16);
    uint2 R0 = lds64(tab + 32);
    uint32_t p1 = min(tab + 48u, tabLast);
    uint4 P1 = lds128(p1);
    float f0 = ldsF(wb + P0.x), f1 = ldsF(wb + P0.y), f2 = ldsF(wb + P0.z);
    float thr0 = __uint_as_float(P0.w);
#pragma unroll 2
    for (int t = 0; t < nT; t++)
    {
        // stage A: offsets of tree t+2 (past the end the last record is re-read)
        const uint32_t p2 = min(p1 + 48u, tabLast);
        const uint4 P2 = lds128(p2);
        // stage B: features and the rest of the record of tree t+1
        const float g0 = ldsF(wb + P1.x), g1 = ldsF(wb + P1.y), g2 = ldsF(wb + P1.z);
        const uint4 Q1 = lds128(p1 + 16);
        const uint2 R1 = lds64(p1 + 32);
        // stage C: decide tree t
        float leaf;
        if (f0 < thr0) leaf = (f1 < __uint_as_float(Q0.x)) ? __uint_as_float(Q0.z) : __uint_as_float(Q0.w);
        else leaf = (f2 < __uint_as_float(Q0.y)) ? __uint_as_float(R0.x) : __uint_as_float(R0.y);
        if (alive)
        {
            h += leaf;
            nEval++;
            if (h <= cascThr) alive = false;
        }
        if (__ballot_sync(FULLMASK, alive) == 0) return 0u;
        thr0 = __uint_as_float(P1.w); Q0 = Q1; R0 = R1; f0 = g0; f1 = g1; f2 = g2;
        P1 = P2; p1 = p2;
    }
    return __ballot_sync(FULLMASK, alive);
}

// Trees [T0, T1) of the table's head (CascTileArgs::head, at least 64 trees): the records are kernel parameters, i.e. constant
// bank operands of the compare / select / add instructions themselves -- no record loads, no table pointer, and the loop is
// fully unrolled so every offset is an immediate.  A dead window is marked by a score of -inf (leaf values are finite --
// Model::validate -- so -inf + leaf stays -inf and the lane can never come back), which replaces the alive flag and its
// bookkeeping: a step is 3 adds + 3 LDS + 3 compares + 2 selects + the score update.  Bit-identical scores for survivors.
template <int T0, int T1>
__device__ __forceinline__ unsigned ctSegHead(const CascTileArgs& a, uint32_t wb, bool valid, float& h, unsigned& nEval)
{
    // A dead window carries a NaN score: NaN + leaf stays NaN and (NaN <= cascThr) is false, so the death test fires exactly once
    // per lane and records the tree count with a predicated move -- no alive flag, no per-tree counter.  The leaf is not selected
    // and then added: the three compares produce the four path predicates directly (setp with two destinations and a predicate
    // input) and each guards an add whose operand is the leaf in the constant bank: 3 setp + 4 add per tree, no select, no
    // constant load.  Scores of surviving windows are the same sequential float sums as before, bit for bit.
    const float cascThr = a.cascThr;
    float s = valid ? h : __int_as_float(0x7fc00000);
    unsigned cnt = 0; // trees this lane evaluated in the segment; set when it dies
    float f0 = ldsF(wb + a.head[T0].off[0]), f1 = ldsF(wb + a.head[T0].off[1]), f2 = ldsF(wb + a.head[T0].off[2]);
#pragma unroll
    for (int t = T0; t < T1; t++)
    {
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        if (t + 1 < T1) { g0 = ldsF(wb + a.head[t + 1].off[0]); g1 = ldsF(wb + a.head[t + 1].off[1]); g2 = ldsF(wb + a.head[t + 1].off[2]); }
        asm volatile("{\n\t.reg .pred p0, p1, p2, p3, p4, p5, pd;\n\t"
                     "setp.lt.f32 p0|p1, %2, %5;\n\t"
                     "setp.lt.and.f32 p2|p3, %3, %6, p0;\n\t"
                     "setp.lt.and.f32 p4|p5, %4, %7, p1;\n\t"
                     "@p2 add.f32 %0, %0, %8;\n\t"
                     "@p3 add.f32 %0, %0, %9;\n\t"
                     "@p4 add.f32 %0, %0, %10;\n\t"
                     "@p5 add.f32 %0, %0, %11;\n\t"
                     "setp.le.f32 pd, %0, %12;\n\t"
                     "@pd mov.b32 %0, 0x7fc00000;\n\t"
                     "@pd mov.u32 %1, %13;\n\t}"
                     : "+f"(s), "+r"(cnt)
                     : "f"(f0), "f"(f1), "f"(f2), "f"(a.head[t].thr[0]), "f"(a.head[t].thr[1]), "f"(a.head[t].thr[2]),
                       "f"(a.head[t].leaf[0]), "f"(a.head[t].leaf[1]), "f"(a.head[t].leaf[2]), "f"(a.head[t].leaf[3]), "f"(cascThr),
                       "r"((unsigned)(t - T0 + 1)));
        if (((t - T0) & 7) == 7 && t + 1 < T1)
            if (__ballot_sync(FULLMASK, s == s) == 0) { nEval += cnt; return 0u; }
        f0 = g0; f1 = g1; f2 = g2;
    }
    const bool alive = (s == s);
    nEval += alive ? (unsigned)(T1 - T0) : cnt;
    h = s;
    return __ballot_sync(FULLMASK, alive);
}

// The same trees on ONE window per warp, lanes = trees: every lane evaluates its own tree of a group of 32 (record and
// features are independent of the running score), then the 32 leaf values are added in tree order -- the reference's
// sequential float sum, identical in every lane -- until the score drops to cascThr.  32 trees cost one round of loads
// instead of 32 dependent steps: the form for the few windows that are still alive deep in the cascade (hits walk every
// tree), where 32-window batches would leave the block waiting on one nearly empty warp.
__device__ __forceinline__ bool ctSparse(uint32_t tab, int nT, float cascThr, uint32_t wb, int lane, float& h, unsigned& nEval)
{
    for (int s = 0; s < nT; s += 32)
    {
        const uint32_t p = tab + 48u * (uint32_t)min(s + lane, nT - 1);
        const uint4 P = lds128(p), Q = lds128(p + 16);
        const uint2 R = lds64(p + 32);
        const float f0 = ldsF(wb + P.x), f1 = ldsF(wb + P.y), f2 = ldsF(wb + P.z);
        float leaf;
        if (f0 < __uint_as_float(P.w)) leaf = (f1 < __uint_as_float(Q.x)) ? __uint_as_float(Q.z) : __uint_as_float(Q.w);
        else leaf = (f2 < __uint_as_float(Q.y)) ? __uint_as_float(R.x) : __uint_as_float(R.y);
        const int cnt = min(32, nT - s);
#pragma unroll 8
        for (int j = 0; j < cnt; j++)
        {
            h += __shfl_sync(FULLMASK, leaf, j);
            if (h <= cascThr)
            {
                if (lane == 0) nEval += (unsigned)(s + j + 1);
                return false;
            }
        }
    }
    if (lane == 0) nEval += (unsigned)nT;
    return true;
}

// THREADS = 512 with two blocks per SM (112 KB of shared memory each) or 384 with three (75 KB each: smaller tiles, more halo,
// but a third tile in flight per SM while the others wait at their level barriers)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 3) k_cascade_tile(const __grid_constant__ CascTileArgs a)
{
    constexpr int kCtThreads = THREADS;
    extern __shared__ __align__(128) uint8_t ctSm[];
    uint8_t* tile = ctSm;
    uint32_t* tabRes = reinterpret_cast<uint32_t*>(ctSm + a.tileBytes);
    uint32_t* tabRing = tabRes + kCtChunk * kCtRecWords;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tabRing + 2 * kCtChunk * kCtRecWords); // [0] tile, [1] resident table, [2], [3] ring slots
    volatile int* si = reinterpret_cast<volatile int*>(bars + 4);
    float* ls = reinterpret_cast<float*>(const_cast<int*>(si) + 32);                    // [2][listCap] scores of the survivors
    uint16_t* lw = reinterpret_cast<uint16_t*>(ls + 2 * a.listCap);                     // [2][listCap] window (c << 8 | r) inside the tile
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    constexpr int nWarps = kCtThreads / 32;
    const long long total = (long long)a.tilesPerFrame * a.n;
    const int nTrees = a.nTrees;
    const int nHead = a.headLevels;
    int nLevels = 1;
    while (ctSegEnd(nLevels - 1, nHead) < nTrees) nLevels++;
    const float cascThr = a.cascThr;
    const int rowBatches = a.Wr >> 5;
    const uint32_t tileAddr = smemU32(tile);
    unsigned nEval = 0;
    unsigned long long nWin = 0;

    // task -> (frame, scale, tile) -> block scalars; issues the tile copy.  Thread 0 only.
    auto startTile = [&](long long task) {
        si[kSiTask] = task < total ? 1 : 0;
        if (task >= total) return;
        const int f = (int)(task / a.tilesPerFrame);
        const int tk = (int)(task - (long long)f * a.tilesPerFrame);
        int lo = 0, hi = a.nScales - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.scales[mid].tile0 <= tk) lo = mid; else hi = mid - 1; }
        const CascTileScale S = a.scales[lo];
        const int tl = tk - S.tile0, tx = tl / S.nTy, ty = tl - tx * S.nTy;
        const int c0 = tx * a.Wc, r0 = ty * a.Wr;
        si[kSiFrame] = f; si[kSiScale] = lo; si[kSiC0] = c0; si[kSiR0] = r0;
        si[kSiNc] = min(a.Wc, S.width1 - c0); si[kSiNr] = min(a.Wr, S.height1 - r0); si[kSiScaleIdx] = S.scaleIdx;
        si[kSiCnt] = 0; si[kSiCnt + 1] = 0; si[kSiCnt + 2] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the block's generic-proxy reads of the old tile are ordered before the copy engine's writes
        mbarExpectTx(&bars[0], (unsigned)a.boxBytes);
        tmaLoad4d(tile, a.maps + lo, &bars[0], r0 * a.step, c0 * a.step, 0, a.frame0 + f);
    };
    auto prefetchTile = [&](long long task) {
        if (task >= total) return;
        const int f = (int)(task / a.tilesPerFrame);
        const int tk = (int)(task - (long long)f * a.tilesPerFrame);
        int lo = 0, hi = a.nScales - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.scales[mid].tile0 <= tk) lo = mid; else hi = mid - 1; }
        const int tl = tk - a.scales[lo].tile0, nTy = a.scales[lo].nTy, tx = tl / nTy, ty = tl - tx * nTy;
        tmaPrefetchL2(a.maps + lo, ty * a.Wr * a.step, tx * a.Wc * a.step, 0, a.frame0 + f);
    };

    if (tid == 0)
    {
        mbarInit(&bars[0], 1); mbarInit(&bars[1], 1); mbarInit(&bars[2], 1); mbarInit(&bars[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
    {
        const unsigned resBytes = (unsigned)(min(nTrees, kCtChunk) * kCtRecWords * 4);
        mbarExpectTx(&bars[1], resBytes);
        bulkLoad(tabRes, a.tab, resBytes, &bars[1]);
        startTile((long long)atomicAdd(a.taskCounter, 1ull));
    }
    __syncthreads();
    mbarWait(&bars[1], 0);
    unsigned tilePhase = 0, ringPhase0 = 0, ringPhase1 = 0;

    while (si[kSiTask])
    {
        long long nextTask = 0;
        if (tid == 0)
        {   // the tile after this one: claim it now and pull it towards L2 while this one is decided
            nextTask = (long long)atomicAdd(a.taskCounter, 1ull);
            prefetchTile(nextTask);
        }
        const int frame = si[kSiFrame], c0 = si[kSiC0], r0 = si[kSiR0], nc = si[kSiNc], nr = si[kSiNr], scaleIdx = si[kSiScaleIdx];
        if (tid == 0) nWin += (unsigned long long)nc * nr;
        if (wib == 0) mbarWait(&bars[0], tilePhase); // one warp polls; the others sleep at the block barrier instead of spinning
        __syncthreads();
        mbarWait(&bars[0], tilePhase);               // completed: every thread observes the phase itself (acquires the copy engine's writes)
        tilePhase ^= 1;
        for (int l = 0; l < nLevels; l++)
        {
            const int tBeg = l == 0 ? 0 : ctSegEnd(l - 1, nHead), tEnd = min(ctSegEnd(l, nHead), nTrees);
            const int slot = (l - nHead) & 1; // ring slot of a streamed level
            uint32_t tab;
            if (l < nHead) tab = smemU32(tabRes) + 48u * (uint32_t)tBeg;
            else
            {   // this level's chunk was requested one level ago
                if (slot) { mbarWait(&bars[3], ringPhase1); ringPhase1 ^= 1; }
                else { mbarWait(&bars[2], ringPhase0); ringPhase0 ^= 1; }
                tab = smemU32(tabRing + slot * kCtChunk * kCtRecWords);
            }
            const int nIn = l == 0 ? nc * rowBatches * 32 : si[kSiCnt + l % 3];
            if (l >= nHead && nIn > 0 && nIn <= a.exportMax)
            {   // A handful of windows is still alive past tree 64 (hits walk every tree).  Finishing them here would keep the
                // tile -- and half an SM -- waiting on one nearly empty warp for up to nTrees sequential steps; hand them to
                // k_cascade_tail (one window per warp, 32 trees per step) and move on to the next tile.
                if (tid == 0)
                {
                    int base = -1, old = *reinterpret_cast<volatile int*>(a.tailCount);
                    while (old + nIn <= a.tailCap)
                    {
                        const int prev = atomicCAS(a.tailCount, old, old + nIn);
                        if (prev == old) { base = old; break; }
                        old = prev;
                    }
                    si[kSiExport] = base;
                }
                __syncthreads();
                const int base = si[kSiExport];
                if (base >= 0)
                {
                    const float* lsE = ls + (l & 1) * a.listCap;
                    const uint16_t* lwE = lw + (l & 1) * a.listCap;
                    const int sl = si[kSiScale];
                    for (int i = tid; i < nIn; i += kCtThreads)
                    {
                        const uint32_t w = lwE[i];
                        a.tail[base + i] = make_int4(frame | (sl << 24), (c0 + (int)(w >> 8)) | ((r0 + (int)(w & 0xffu)) << 16), __float_as_int(lsE[i]), tBeg);
                    }
                    break;
                }
            }
            if (tid == 0)
            {
                si[kSiCnt + (l + 2) % 3] = 0; // the counter level l+1 appends to (its readers passed the previous barrier)
                if (l >= nHead - 1 && l + 1 < nLevels && nIn > 0)
                {   // next level's records -> the ring slot level l-1 has finished with
                    const int t0 = ctSegEnd(l, nHead), cnt = min(kCtChunk, nTrees - t0);
                    const int nslot = (l + 1 - nHead) & 1;
                    uint64_t* bar = &bars[2 + nslot];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbarExpectTx(bar, (unsigned)(cnt * kCtRecWords * 4));
                    bulkLoad(tabRing + nslot * kCtChunk * kCtRecWords, a.tab + (size_t)t0 * kCtRecWords, (unsigned)(cnt * kCtRecWords * 4), bar);
                }
            }
            if (nIn == 0) break;
            const float* lsIn = ls + (l & 1) * a.listCap;
            const uint16_t* lwIn = lw + (l & 1) * a.listCap;
            float* lsOut = ls + ((l + 1) & 1) * a.listCap;
            uint16_t* lwOut = lw + ((l + 1) & 1) * a.listCap;
            volatile int* cntOut = si + kSiCnt + (l + 1) % 3;
            const bool last = tEnd >= nTrees;
            if (l >= nHead && nIn <= a.sparseMax)
            {   // few survivors deep in the cascade: one window per warp, lanes = trees
                for (int i = wib; i < nIn; i += nWarps)
                {
                    const uint32_t win = lwIn[i];
                    float h = lsIn[i];
                    const uint32_t wb = tileAddr + ((win >> 8) * (uint32_t)a.BY + (win & 0xffu)) * (uint32_t)(a.step * 4);
                    if (!ctSparse(tab, tEnd - tBeg, cascThr, wb, lane, h, nEval) || lane != 0) continue;
                    if (last)
                    {
                        if (h > cascThr)
                        {
                            const int idx = atomicAdd(a.hitCount + frame, 1);
                            if (idx < a.cap) a.hits[(size_t)frame * a.cap + idx] = make_int4(scaleIdx, c0 + (int)(win >> 8), r0 + (int)(win & 0xffu), __float_as_int(h));
                        }
                    }
                    else
                    {
                        const int pos = atomicAdd(const_cast<int*>(cntOut), 1);
                        lwOut[pos] = (uint16_t)win; lsOut[pos] = h;
                    }
                }
            }
            else
            for (int b = wib; b * 32 < nIn; b += nWarps)
            {
                bool valid;
                uint32_t win;
                float h = 0.f;
                if (l == 0)
                {
                    // (the usual tile is 32 window rows tall: one batch per window column, no division)
                    const int c = rowBatches == 1 ? b : b / rowBatches, rr = rowBatches == 1 ? lane : (b - c * rowBatches) * 32 + lane;
                    valid = rr < nr;
                    win = (uint32_t)(c << 8) | (uint32_t)rr;
                }
                else
                {
                    const int i = b * 32 + lane;
                    valid = i < nIn;
                    win = valid ? lwIn[i] : 0u;
                    if (valid) h = lsIn[i];
                }
                if (!valid) win = 0u;
                const uint32_t wb = tileAddr + ((win >> 8) * (uint32_t)a.BY + (win & 0xffu)) * (uint32_t)(a.step * 4);
                unsigned surv;
                if (l < nHead && a.headTrees)
                {
                    switch (l)
                    {
                        case 0: surv = ctSegHead<0, 4>(a, wb, valid, h, nEval); break;
                        case 1: surv = ctSegHead<4, 8>(a, wb, valid, h, nEval); break;
                        case 2: surv = nHead == 3 ? ctSegHead<8, 64>(a, wb, valid, h, nEval) : ctSegHead<8, 16>(a, wb, valid, h, nEval); break;
                        case 3: surv = ctSegHead<16, 32>(a, wb, valid, h, nEval); break;
                        default: surv = ctSegHead<32, 64>(a, wb, valid, h, nEval); break;
                    }
                }
                else surv = ctSegment(tab, tEnd - tBeg, cascThr, wb, valid, h, nEval);
                if (!surv) continue;
                const bool mine = (surv >> lane) & 1u;
                if (last)
                {
                    if (mine && h > cascThr)
                    {
                        const int idx = atomicAdd(a.hitCount + frame, 1);
                        if (idx < a.cap) a.hits[(size_t)frame * a.cap + idx] = make_int4(scaleIdx, c0 + (int)(win >> 8), r0 + (int)(win & 0xffu), __float_as_int(h));
                    }
                }
                else
                {
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(const_cast<int*>(cntOut), __popc(surv));
                    pos = __shfl_sync(FULLMASK, pos, 0) + __popc(surv & ((1u << lane) - 1u));
                    if (mine) { lwOut[pos] = (uint16_t)win; lsOut[pos] = h; }
                }
            }
            __syncthreads();
        }
        __syncthreads(); // every warp is done with the tile and the block scalars
        if (tid == 0) startTile(nextTask);
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nEval += __shfl_down_sync(FULLMASK, nEval, o);
    if (lane == 0 && nEval) atomicAdd(a.stats, (unsigned long long)nEval);
    if (tid == 0 && nWin) atomicAdd(a.stats + 1, nWin);
}

// ------------------------------------------------------------------------------------------------------------------
// k_cascade_tail: finishes the windows k_cascade_tile handed over (CascTileArgs::tail).  One window per warp, lanes =
// trees: every lane fetches the record of its own tree of a group of 32 and gathers its three features from the
// pyramid in global memory (a window's 11-16 KB footprint stays in L1 / L2 over the 64 groups of a 2048-tree model);
// records and features do not depend on the running score, so the next group's loads are in flight while this
// group's 32 leaf values are added in tree order -- the reference's sequential float sum -- and the first tree at
// which the score drops to cascThr is found with one ballot.  A hit costs 64 such steps instead of 2048 dependent ones.
// ------------------------------------------------------------------------------------------------------------------
struct TailSlot // one window in flight in a warp of k_cascade_tail
{
    bool act;
    const float* chns; // the window's origin in its pyramid plane 0
    unsigned P, planeStride;
    int frame, c, r, scaleIdx, s; // s: first tree of the group whose leaf values are in `leaf`
    float h, leaf;
};

__global__ void __launch_bounds__(256, 4) k_cascade_tail(CascTailArgs a)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int n = min(*a.tailCount, a.tailCap);
    const int shShift = __ffs(a.shrink) - 1;
    const uint4* __restrict__ tab = reinterpret_cast<const uint4*>(a.tab);
    unsigned long long nEval = 0;
    int nextIdx = gw;
    auto leafOf = [&](const TailSlot& z, int t) {   // tree t of this lane: root and both children gathered together (acfDetect1.cpp:100-138, depth 2)
        const uint4* rec = tab + (size_t)min(t, a.nTrees - 1) * 4;
        const uint4 n0 = __ldg(rec), n1 = __ldg(rec + 1), n2 = __ldg(rec + 2), lf = __ldg(rec + 3);
        const float f0 = __ldg(z.chns + (n0.x * z.planeStride + n0.y * z.P + n0.z));
        const float f1 = __ldg(z.chns + (n1.x * z.planeStride + n1.y * z.P + n1.z));
        const float f2 = __ldg(z.chns + (n2.x * z.planeStride + n2.y * z.P + n2.z));
        if (f0 < __uint_as_float(n0.w)) return (f1 < __uint_as_float(n1.w)) ? __uint_as_float(lf.x) : __uint_as_float(lf.y);
        return (f2 < __uint_as_float(n2.w)) ? __uint_as_float(lf.z) : __uint_as_float(lf.w);
    };
    auto fetch = [&](TailSlot& z) {   // next window of this warp's share of the list (warp uniform)
        z.act = nextIdx < n;
        if (!z.act) { z.leaf = 0.f; z.h = 0.f; return; }
        const int4 e = a.tail[nextIdx];
        nextIdx += nw;
        z.frame = e.x & 0xffffff; z.c = e.y & 0xffff; z.r = (unsigned)e.y >> 16;
        const CascScale S = a.scales[(unsigned)e.x >> 24];
        z.P = (unsigned)S.P; z.planeStride = (unsigned)S.planeStride; z.scaleIdx = S.scaleIdx;
        z.chns = a.pyr + z.frame * a.frameStride + S.off + (size_t)((z.c * a.stride) >> shShift) * S.P + ((z.r * a.stride) >> shShift);
        z.h = __int_as_float(e.z); z.s = e.w;
        z.leaf = leafOf(z, z.s + lane);
    };
    // after the ripple: v = score after this lane's tree; retire the window (dead, or past the last tree) or move it on
    auto settle = [&](TailSlot& z, float v, float leafNext) {
        if (!z.act) return;
        const int cnt = min(32, a.nTrees - z.s);
        const unsigned dead = __ballot_sync(FULLMASK, lane < cnt && v <= a.cascThr);
        if (dead) { nEval += (unsigned)__ffs(dead); fetch(z); return; }
        nEval += (unsigned)cnt;
        z.h = __shfl_sync(FULLMASK, v, cnt - 1);
        z.s += 32;
        if (z.s >= a.nTrees)
        {
            if (lane == 0 && z.h > a.cascThr)
            {
                const int idx = atomicAdd(a.hitCount + z.frame, 1);
                if (idx < a.cap) a.hits[(size_t)z.frame * a.cap + idx] = make_int4(z.scaleIdx, z.c, z.r, __float_as_int(z.h));
            }
            fetch(z);
            return;
        }
        z.leaf = leafNext;
    };
    // Two windows are in flight per warp: their ripples are independent dependency chains, so one hides the other's
    // shuffle latency, and a window that dies is replaced from the list while its neighbour carries on.
    TailSlot z0, z1;
    fetch(z0); fetch(z1);
    while (z0.act || z1.act)
    {
        // next group's leaf values: records and features do not depend on the running score, in flight during the ripple
        const float ln0 = (z0.act && z0.s + 32 < a.nTrees) ? leafOf(z0, z0.s + 32 + lane) : 0.f;
        const float ln1 = (z1.act && z1.s + 32 < a.nTrees) ? leafOf(z1, z1.s + 32 + lane) : 0.f;
        // The reference's sequential sum h += leaf(t) as a ripple through the lanes: lane j ends up with the score after its
        // own tree, ((h + leaf_0) + leaf_1) + ... + leaf_j.  Every pass hands each lane its lower neighbour's value (shfl.up;
        // lane 0 keeps h + leaf_0: the shuffle's range predicate guards the add), so after pass k lanes 0..k hold their final
        // values and keep recomputing the same numbers -- two instructions per tree instead of a broadcast, an add and a select.
        float v0 = z0.h + z0.leaf, v1 = z1.h + z1.leaf;
#pragma unroll
        for (int j = 1; j < 32; j++)
            asm volatile("{\n\t.reg .pred p, q;\n\t.reg .f32 t, u;\n\t"
                         "shfl.sync.up.b32 t|p, %0, 1, 0, 0xffffffff;\n\t"
                         "shfl.sync.up.b32 u|q, %1, 1, 0, 0xffffffff;\n\t"
                         "@p add.f32 %0, t, %2;\n\t"
                         "@q add.f32 %1, u, %3;\n\t}"
                         : "+f"(v0), "+f"(v1) : "f"(z0.leaf), "f"(z1.leaf));
        settle(z0, v0, ln0);
        settle(z1, v1, ln1);
    }
    if (lane == 0 && nEval) atomicAdd(a.stats, nEval);
}

// ------------------------------------------------------------------------------------------------------------------
// k_cascade_tail_win: the same hand-over, each window finished on its own shared-memory footprint.
// Why: k_cascade_tail's three feature gathers per tree are 32 different 128-byte lines per instruction (lanes = trees:
// addresses all over the window) -- ncu shows the L1 data pipe at 91 % of its wavefront peak and 26 % issue-active --
// and its ripple adds a shuffle wavefront per tree.  Re-staging whole tiles does not pay either: a hand-over holds 1-2
// windows on average (measured: the block then waits at its barrier for the one warp that has a window).  So every WARP
// is on its own here: it takes a window, brings exactly that window's footprint (modelHt/shrink rows x modelWd/shrink
// columns x all channels, 11 KB for the face models) into its shared-memory slot with one 4-D cp.async.bulk.tensor
// behind its own mbarrier, and walks the remaining trees 32 at a time, lanes = trees:
//   * the three features of a tree are shared-memory loads (a few bank-conflict wavefronts instead of 32 lines),
//   * the records come from a structure-of-arrays table with window-local offsets (a warp's 32 trees: 32 sectors),
//   * the 32 leaf values go through a 128-byte shared-memory row; every lane then adds the row in tree order in
//     registers -- the reference's sequential float sum, acfDetect1.cpp:100-138 -- tracking the running minimum: one
//     store, eight broadcast loads and 48 arithmetic instructions per 32 trees, no shuffles.  The exact tree of death
//     (needed for the trees-evaluated statistic only) is searched when the minimum says the window died.
// No block-wide synchronisation anywhere: sixteen warps per SM, each at its own window and tree.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTwWarps = 8;
__global__ void __launch_bounds__(32 * kTwWarps, 2) k_cascade_tail_win(const __grid_constant__ CascTailWinArgs a)
{
    extern __shared__ __align__(128) uint8_t twSm[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int footStride = (a.footBytes + 127) & ~127;
    uint8_t* foot = twSm + wib * footStride;
    float* row = reinterpret_cast<float*>(twSm + kTwWarps * footStride) + wib * 32;
    uint64_t* bar = reinterpret_cast<uint64_t*>(twSm + kTwWarps * footStride + kTwWarps * 128) + wib;
    const uint32_t footAddr = smemU32(foot), rowAddr = smemU32(row);
    const int gw = (blockIdx.x * blockDim.x + tid) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int n = min(*a.tailCount, a.tailCap);
    const float cascThr = a.cascThr;
    const int nTrees = a.nTrees;
    unsigned long long nEval = 0; // identical in every lane of the warp
    if (lane == 0)
    {
        mbarInit(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    for (int i = gw; i < n; i += nw)
    {
        const int4 e = a.tail[i];
        const int frame = e.x & 0xffffff, sl = (unsigned)e.x >> 24, c = e.y & 0xffff, r = (unsigned)e.y >> 16;
        if (lane == 0)
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the previous window's reads of the slot come before the copy engine's writes
            mbarExpectTx(bar, (unsigned)a.footBytes);
            // a bulk tensor copy starts at a 16-byte aligned global address: take the rows from the aligned row at or above the window's
            // first one (the box is three rows taller than the window for that)
            tmaLoad4d(foot, a.maps + sl, bar, (r * a.step) & ~3, c * a.step, 0, a.frame0 + frame);
        }
        const uint32_t winAddr = footAddr + 4u * (uint32_t)((r * a.step) & 3);
        float h = __int_as_float(e.z);
        int s = e.w;
        // Software pipeline over groups of 32 trees: while group g is summed, the features of group g + 1 and the records of
        // group g + 2 are in flight (record -> feature -> decision is a chain of two dependent loads)
        int t = min(s + lane, nTrees - 1);
        uint4 A0 = __ldg(a.tabA + t), B0 = __ldg(a.tabB + t);
        uint2 C0 = __ldg(a.tabC + t);
        t = min(s + 32 + lane, nTrees - 1);
        uint4 A1 = __ldg(a.tabA + t), B1 = __ldg(a.tabB + t);
        uint2 C1 = __ldg(a.tabC + t);
        const int scaleIdx = a.scales[sl].scaleIdx;
        mbarWait(bar, phase);
        phase ^= 1;
        float f0 = ldsF(winAddr + A0.x), f1 = ldsF(winAddr + A0.y), f2 = ldsF(winAddr + A0.z);
        bool alive = true;
        for (; s < nTrees; s += 32)
        {
            // features of the next group, records of the one after it
            const float g0 = ldsF(winAddr + A1.x), g1 = ldsF(winAddr + A1.y), g2 = ldsF(winAddr + A1.z);
            t = min(s + 64 + lane, nTrees - 1);
            const uint4 A2 = __ldg(a.tabA + t), B2 = __ldg(a.tabB + t);
            const uint2 C2 = __ldg(a.tabC + t);
            const float thr0 = __uint_as_float(A0.w), thr1 = __uint_as_float(B0.x), thr2 = __uint_as_float(B0.y);
            const float l0 = __uint_as_float(B0.z), l1 = __uint_as_float(B0.w), l2 = __uint_as_float(C0.x), l3 = __uint_as_float(C0.y);
            const float leaf = (f0 < thr0) ? ((f1 < thr1) ? l0 : l1) : ((f2 < thr2) ? l2 : l3);
            A0 = A1; B0 = B1; C0 = C1; A1 = A2; B1 = B2; C1 = C2;
            f0 = g0; f1 = g1; f2 = g2;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(rowAddr + 4u * lane), "f"(leaf) : "memory");
            __syncwarp();
            const int cnt = min(32, nTrees - s);
            float L[32];
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                const uint4 v = lds128(rowAddr + 16u * q);
                L[4 * q] = __uint_as_float(v.x); L[4 * q + 1] = __uint_as_float(v.y); L[4 * q + 2] = __uint_as_float(v.z); L[4 * q + 3] = __uint_as_float(v.w);
            }
            float hs = h, mn = __int_as_float(0x7f800000);
            if (cnt == 32)
            {
#pragma unroll
                for (int j = 0; j < 32; j++) { hs += L[j]; mn = fminf(mn, hs); }
            }
            else
            {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (j < cnt) { hs += L[j]; mn = fminf(mn, hs); }
            }
            if (mn <= cascThr)
            {   // died inside this group: find the tree (the statistic counts trees evaluated, the killing one included)
                float h2 = h;
                int j = 0;
                for (; j < cnt; j++) { h2 += ldsF(rowAddr + 4u * j); if (h2 <= cascThr) break; }
                nEval += (unsigned)(j + 1);
                alive = false;
            }
            else { nEval += (unsigned)cnt; h = hs; }
            __syncwarp(); // the row is free for the next group
            if (!alive) break;
        }
        if (alive && lane == 0 && h > cascThr)
        {
            const int idx = atomicAdd(a.hitCount + frame, 1);
            if (idx < a.cap) a.hits[(size_t)frame * a.cap + idx] = make_int4(scaleIdx, c, r, __float_as_int(h));
        }
        __syncwarp(); // every lane is done with the footprint
    }
    if (lane == 0 && nEval) atomicAdd(a.stats, nEval);
}

int cascTailWinSmem(int footBytes) { return kTwWarps * ((footBytes + 127) & ~127) + kTwWarps * 128 + kTwWarps * 8; }

void launchCascadeTailWin(const CascTailWinArgs& a, cudaStream_t s)
{
    const int smem = cascTailWinSmem(a.footBytes);
    cudaFuncSetAttribute(k_cascade_tail_win, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_cascade_tail_win<<<148 * 2, 32 * kTwWarps, smem, s>>>(a);
}

void launchCascadeTail(const CascTailArgs& a, cudaStream_t s)
{
    k_cascade_tail<<<148 * 4, 256, 0, s>>>(a);
}

// ---- host side ---------------------------------------------------------------------------------------------------

// Tile geometry for a model: the window grid Wc x Wr (Wr a multiple of 32: a fresh batch is 32 consecutive rows of
// one window column) whose footprint + survivor lists + tables fit half an SM's shared memory with the least halo
// amplification.  mW / mH: window size in channel pixels along x / y; step = stride / shrink.
bool cascTileGeometry(int mH, int mW, int nChns, int step, int blocksPerSm, CascTileGeom& g)
{
    const int budget = blocksPerSm >= 3 ? 75 * 1024 : 112 * 1024; // two (or three) blocks per SM
    double best = 1e30;
    bool ok = false;
    for (int Wr = 32; Wr <= 128; Wr += 32)
    {
        const int BY = (((Wr - 1) * step + mH) + 3) & ~3;
        if (BY > 256) break;
        for (int Wc = 1; Wc <= 128; Wc++)
        {
            const int BX = (Wc - 1) * step + mW;
            if (BX > 256) break;
            const int tileBytes = (BX * BY * nChns * 4 + 127) & ~127;
            const int listCap = (Wc * Wr + 1) & ~1;
            const int bytes = tileBytes + kCtFixedBytes + 2 * listCap * 6;
            if (bytes > budget) break;
            const double amp = (double)BX * BY / ((double)Wc * Wr);
            if (amp < best) { best = amp; g.Wc = Wc; g.Wr = Wr; g.BX = BX; g.BY = BY; g.tileBytes = tileBytes; g.boxBytes = BX * BY * nChns * 4; g.listCap = listCap; g.smemBytes = bytes; ok = true; }
        }
    }
    g.step = step; g.nChns = nChns;
    return ok;
}

int cascTileRecWords() { return kCtRecWords; }

void launchCascadeTile(const CascTileArgs& a, cudaStream_t s)
{
    const long long tiles = (long long)a.tilesPerFrame * a.n;
    if (tiles <= 0) return;
    if (a.blocksPerSm >= 3)
    {
        cudaFuncSetAttribute(k_cascade_tile<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75 * 1024); // per device, cheap
        k_cascade_tile<384><<<(int)std::min<long long>(tiles, 148 * 3), 384, a.smemBytes, s>>>(a);
        return;
    }
    cudaFuncSetAttribute(k_cascade_tile<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    const int grid = (int)std::min<long long>(tiles, 148 * (a.blocksPerSm == 1 ? 1 : 2));
    k_cascade_tile<512><<<grid, 512, a.smemBytes, s>>>(a);
}

} // namespace acfb
