// post.cu -- k_post: the caller-side tail of Detector::operator()(Pyramid) on the device, one thread block per frame:
//   * order the raw hits the way the reference's loops produce them (scale-major, then window column, then row;
//     ACF.cpp:326-330, acfDetect1.cpp:86-97) and rescale them to boxes (ACF.cpp:302-311: int truncation of the shifted
//     corner, cvRound of modelDs / scale, x <-> y swap),
//   * bbNms type "max" / "maxg" (bbNms.cpp:111-192: score-descending order, overlap = intersection / union or / min area
//     in double, a box suppresses every later box it overlaps by more than pNms.overlap; "maxg": only boxes still alive
//     suppress),
//   * ObjectDetector::prune (ObjectDetector.cpp:28-44): at most maxDet boxes, cut after the first one whose score is
//     below ratio * best.
// Everything is integer / double arithmetic with the reference's statements, so the boxes equal the host tail's
// (engine.cu: boxOf, nmsAndPrune) bit for bit; ties in the score order resolve by the reference order, like the host's
// stable sort.  What leaves the device is at most maxDet boxes per frame instead of every raw hit, and the records are
// where an NCCL gather can take them from (dist.cu) without a host round trip.
#include "kernels.cuh"
#include <cstdio>

namespace acfb
{

constexpr int kPostThreads = 256;

__device__ __forceinline__ uint32_t orderedScore(float s)
{   // float -> unsigned that sorts ascending like the float (negative zero below positive zero: scores are sums, the
    // host's comparison `a.score > b.score` treats -0 == +0, and so does the tie-break below because equal floats other
    // than the two zeros have equal bits)
    const uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(kPostThreads) k_post(PostArgs a)
{
    extern __shared__ __align__(16) unsigned long long postSm[];
    unsigned long long* key = postSm;                                   // [cap2] sort keys: ~score | reference order
    int4* box = reinterpret_cast<int4*>(postSm + a.cap2);               // [cap2] x, y, w, h by sorted position
    float* sc = reinterpret_cast<float*>(box + a.cap2);                 // [cap2] score by sorted position
    unsigned char* kp = reinterpret_cast<unsigned char*>(sc + a.cap2);  // [cap2] still alive
    __shared__ int sKeep[64];
    __shared__ int sN;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int raw = a.hitCount[f];
    if (raw > a.cap2 || raw > a.hitCap)
    {   // more hits than the shared-memory sort holds (or the hit buffer overflowed): the host tail takes this batch
        if (tid == 0) { a.detCount[f] = -1; atomicExch(a.fallback, 1); }
        return;
    }
    const int4* hits = a.hits + (size_t)f * a.hitCap;
    // smallest power of two >= raw for the bitonic network; padding keys sort last
    int n2 = 1;
    while (n2 < raw) n2 <<= 1;
    for (int i = tid; i < n2; i += kPostThreads)
    {
        unsigned long long k = ~0ull;
        if (i < raw)
        {
            const int4 h = hits[i];
            // descending score first, then the reference's order: scale, window column c, row r
            const uint32_t ord = ((uint32_t)h.x << 26) | ((uint32_t)h.y << 13) | (uint32_t)h.z;
            float s = __int_as_float(h.w);
            if (s == 0.0f) s = 0.0f; // -0 -> +0: compares equal on the host
            k = ((unsigned long long)(~orderedScore(s)) << 32) | ord;
        }
        key[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1)
        {
            for (int i = tid; i < (n2 >> 1); i += kPostThreads)
            {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long x = key[lo], y = key[hi];
                if ((x > y) == up) { key[lo] = y; key[hi] = x; }
            }
            __syncthreads();
        }
    // boxes in sorted order (ACF.cpp:302-311)
    for (int i = tid; i < raw; i += kPostThreads)
    {
        const uint32_t ord = (uint32_t)key[i];
        const int s = ord >> 26, c = (ord >> 13) & 0x1fff, r = ord & 0x1fff;
        const PostScale S = a.scales[s];
        int rx = r * a.stride, ry = c * a.stride;
        const int sw = (int)rint(double(a.modelDs_w) / S.scale); // cvRound
        const int sh = (int)rint(double(a.modelDs_h) / S.scale);
        rx = (int)(double(rx + a.shift_w) / S.shw_w);            // int truncation
        ry = (int)(double(ry + a.shift_h) / S.shw_h);
        box[i] = make_int4(ry, rx, sh, sw);
        const uint32_t os = ~(uint32_t)(key[i] >> 32);
        sc[i] = __uint_as_float((os & 0x80000000u) ? (os & 0x7fffffffu) : ~os);
        kp[i] = 1;
    }
    __syncthreads();
    auto overlaps = [&](const int4 bi, const int4 bj) -> bool {
        const int iw = min(bi.x + bi.z, bj.x + bj.z) - max(bi.x, bj.x);
        if (iw <= 0) return false;
        const int ih = min(bi.y + bi.w, bj.y + bj.w) - max(bi.y, bj.y);
        if (ih <= 0) return false;
        double o = (double)(iw * ih);
        const int ias = bi.z * bi.w, jas = bj.z * bj.w;
        const double u = a.ovrUnion ? ((double)(ias + jas) - o) : (double)min(ias, jas);
        o /= u;
        return o > a.overlap;
    };
    const int want = min(a.maxDet, 64);
    if (a.greedy)
    {   // only the first `want` survivors are ever reported (prune), and a box's fate depends on earlier survivors only: walk the
        // score order, let each survivor suppress everything after it, stop once `want` survivors are known
        if (tid == 0) sN = 0;
        __syncthreads();
        int i = 0;
        while (i < raw && sN < want)
        {
            // next box still alive (block-uniform: every thread scans the same bytes)
            while (i < raw && !kp[i]) i++;
            if (i >= raw) break;
            if (tid == 0) sKeep[sN] = i;
            const int4 bi = box[i];
            for (int j = i + 1 + tid; j < raw; j += kPostThreads)
                if (kp[j] && overlaps(bi, box[j])) kp[j] = 0;
            __syncthreads();
            if (tid == 0) sN = sN + 1;
            __syncthreads();
            i++;
        }
    }
    else
    {   // "max": every box, alive or not, suppresses the later boxes it overlaps -> box j survives iff no earlier box overlaps it
        for (int j = tid; j < raw; j += kPostThreads)
        {
            const int4 bj = box[j];
            bool alive = true;
            for (int i = 0; i < j && alive; i++)
                if (overlaps(box[i], bj)) alive = false;
            kp[j] = alive ? 1 : 0;
        }
        __syncthreads();
        if (tid == 0)
        {
            int m = 0;
            for (int i = 0; i < raw && m < want; i++)
                if (kp[i]) sKeep[m++] = i;
            sN = m;
        }
        __syncthreads();
    }
    if (tid == 0)
    {   // ObjectDetector::prune on the survivors (they are in score order)
        int m = sN;
        // size of the full survivor list matters only through min(maxDet, size): when the walk stopped early there are at least `want`
        if (m > 1)
        {
            int cutoff = 1;
            const int lim = min(a.maxDet, m);
            for (int i = 1; i < lim; i++)
            {
                cutoff = i + 1;
                if ((double)sc[sKeep[i]] < ((double)sc[sKeep[0]] * a.pruneRatio)) break;
            }
            m = cutoff;
        }
        a.detCount[f] = m;
        for (int i = 0; i < m; i++)
        {
            const int4 b = box[sKeep[i]];
            PostDet d; d.x = b.x; d.y = b.y; d.w = b.z; d.h = b.w; d.score = sc[sKeep[i]]; d.frame = a.frame0 + f;
            a.dets[(size_t)f * a.maxOut + i] = d;
        }
    }
}

size_t postSmemBytes(int cap2) { return (size_t)cap2 * (8 + 16 + 4 + 1) + 16; }

void launchPost(const PostArgs& a, cudaStream_t s)
{
    const size_t smem = postSmemBytes(a.cap2);
    cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_post<<<a.n, kPostThreads, smem, s>>>(a);
}

} // namespace acfb
