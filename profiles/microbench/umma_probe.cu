// tcgen05 (UMMA) probe for the tensor-core form of the rational-ratio kernel (sm_100a).
//
// Checks, against a host model of the operand layout, the assumptions art_sinc_umma.cu rests on:
//   1. K-major, no-swizzle ("interleave") operands stored as K-planes: plane p holds k = 8p..8p+7 of every
//      row, rows at a 16-byte pitch (SBO = 128 B between 8-row groups, LBO = plane size);
//   2. an operand start address that is only 16-byte aligned (row shift of the A operand);
//   3. an N sub-range: B start + 16*j0, D column offset j0, N < 160;
//   4. accumulation over many k-steps, and the error of fp32 accumulation in TMEM;
//   5. issue rate: cycles per MMA for M=128, N in {64,160,256}, K=16, both operands in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

struct Op { unsigned aOff, bOff, aLbo, aSbo, bLbo, bSbo, N, dCol, acc; };

__device__ __forceinline__ unsigned smem_u32 (const void *p) { return (unsigned) __cvta_generic_to_shared (p); }

__device__ __forceinline__ unsigned long long make_desc (unsigned addr, unsigned lbo, unsigned sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long) ((addr >> 4) & 0x3fff);
    d |= (unsigned long long) ((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long) ((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                                   // descriptor version (Blackwell)
    return d;                                          // layout type 0 = no swizzle, base offset 0
}

__device__ __forceinline__ unsigned make_idesc (int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned) (N >> 3) << 17) | ((unsigned) (M >> 4) << 24);   // f32 acc, bf16 x bf16, K-major
}

__device__ __forceinline__ void mma (unsigned tmem, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc)
{
    asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mbar_wait (unsigned long long *bar, unsigned parity)
{
    for (unsigned spins = 0; spins < (1u << 24); ++spins) {
        unsigned done;
        asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                      : "=r"(done) : "r"(smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    printf ("probe: mbarrier never completed\n");
    __trap ();
}

extern __shared__ __align__ (1024) unsigned char smem[];

__global__ void __launch_bounds__ (128, 1)
probe (const unsigned char *aImg, unsigned aBytes, const unsigned char *bImg, unsigned bBytes,
       const Op *ops, int numOps, float *D, int dCols, int timeIters, int timeN, long long *cycles, int bgStores, int pattern)
{
    __shared__ __align__ (8) unsigned long long bar, bar2;
    __shared__ unsigned tmemBase;
    unsigned char *aS = smem, *bS = smem + ((aBytes + 1023) & ~1023u);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (unsigned i = tid; i < aBytes / 16; i += 128) reinterpret_cast<uint4 *> (aS)[i] = reinterpret_cast<const uint4 *> (aImg)[i];
    for (unsigned i = tid; i < bBytes / 16; i += 128) reinterpret_cast<uint4 *> (bS)[i] = reinterpret_cast<const uint4 *> (bImg)[i];
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> async-proxy (tensor core) reads
    if (tid == 0) {
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar)));
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar2)));
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (&tmemBase)), "r"(512));
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = tmemBase;
    unsigned parity = 0;

    if (tid == 0) {
        for (int i = 0; i < numOps; ++i) {
            const Op o = ops[i];
            mma (tm + o.dCol, make_desc (smem_u32 (aS) + o.aOff, o.aLbo, o.aSbo), make_desc (smem_u32 (bS) + o.bOff, o.bLbo, o.bSbo),
                 make_idesc (128, (int) o.N), o.acc);
        }
        asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar)) : "memory");
    }
    mbar_wait (&bar, parity); parity ^= 1;
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");

    // read back: lane = 32 * (warp % 4) + lane id, 32 columns per load
    for (int c0 = 0; c0 < dCols; c0 += 32) {
        unsigned r[32];
        const unsigned addr = tm + ((unsigned) (warp * 32) << 16) + (unsigned) c0;
        asm volatile ("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                      : "r"(addr));
        asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t) tid * dCols + c0 + j] = __uint_as_float (r[j]);
    }

    // issue-rate measurement: timeIters MMAs of (128 x timeN x 16) back to back on fresh addresses
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (timeIters > 0) {
        long long t0 = 0, t1 = 0;
        __shared__ volatile int stop;
        __shared__ unsigned long long stored;
        if (tid == 0) { stop = 0; stored = 0; }
        __syncthreads ();
        if (warp > 0 && bgStores) {
            // warps 1-3: 16-byte stores into a scratch area behind the operand images, until the MMAs are done
            uint4 *scratch = reinterpret_cast<uint4 *> (bS + ((bBytes + 1023) & ~1023u)) + (tid - 32);
            unsigned long long n = 0;
            uint4 v = make_uint4 (tid, 1, 2, 3);
            while (!stop) {
#pragma unroll
                for (int u = 0; u < 8; ++u) { scratch[(u & 3) * 96] = v; v.x += 1; }
                n += 8;
                for (int w = 1; w < bgStores; ++w) __nanosleep (0);
            }
            atomicAdd (&stored, n * 16);
        }
        if (tid == 0) {
            const unsigned idesc = make_idesc (128, timeN);
            t0 = clock64 ();
            for (int i = 0; i < timeIters; ++i) {
                const unsigned p = (unsigned) (i % 8) * 2;      // walk over 8 k-steps of the images
                const int m5 = i % 5;
                const unsigned dcol = pattern ? (m5 == 0 ? 0u : (m5 < 3 ? 160u : 320u)) : 0u;
                if (pattern == 2 && m5 == 0 && i) {
                    asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar2)) : "memory");
                    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mma (tm + dcol, make_desc (smem_u32 (aS) + p * 136 * 16 + (i & 3) * 16, 136 * 16, 128),
                     make_desc (smem_u32 (bS) + p * 256 * 16, 256 * 16, 128), idesc, 1);
            }
            asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar)) : "memory");
        }
        if (warp == 0) { mbar_wait (&bar, parity); }
        parity ^= 1;
        if (tid == 0) { t1 = clock64 (); cycles[0] = t1 - t0; stop = 1; }
        __syncthreads ();
        if (tid == 0) cycles[1] = (long long) stored;
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (warp == 0)
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}


__device__ __forceinline__ void mma2 (unsigned tmem, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc)
{
    asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// lean issue loop: five MMAs per iteration with precomputed descriptors (variant selects what differs between them)
__global__ void __launch_bounds__ (128, 1)
probe2 (int variant, int iters, unsigned aLbo, int fmt16, long long *cycles)
{
    __shared__ __align__ (8) unsigned long long bar, bar2, bar3;
    __shared__ unsigned tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (unsigned i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4 *> (smem)[i] = make_uint4 (0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar)));
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar2)));
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar3)));
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (&tmemBase)), "r"(512));
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = tmemBase;
    if (tid == 0) {
        const unsigned base = smem_u32 (smem);
        const unsigned aSplit = 48 * 1024, bBase = base + 100 * 1024, bSplit = 2 * 160 * 16;
        unsigned idesc = (1u << 4) | ((unsigned) (160 >> 3) << 17) | ((unsigned) (128 >> 4) << 24);
        if (!fmt16) idesc |= (1u << 7) | (1u << 10);
        // variant 0: all five identical; 1: three accumulators; 2: + operand pattern of the real kernel
        const unsigned d0 = tm, d1 = variant >= 1 ? tm + 160 : tm, d2 = variant >= 1 ? tm + 320 : tm;
        const unsigned a1 = base, a2 = variant >= 2 ? base + aSplit : base;
        const unsigned b1 = bBase, b2 = variant >= 2 ? bBase + bSplit : bBase, b3 = variant >= 2 ? bBase + 2 * bSplit : bBase;
        const unsigned long long A1 = make_desc (a1, aLbo, 128), A2 = make_desc (a2, aLbo, 128);
        const unsigned long long B1 = make_desc (b1, 2560, 128), B2 = make_desc (b2, 2560, 128), B3 = make_desc (b3, 2560, 128);
        const long long t0 = clock64 ();
        for (int it = 0; it < iters; ++it) {
            const unsigned long long sh = (unsigned long long) ((it & 3) + 2 * (it & 7) * (aLbo >> 4));   // row shift + plane pair walk
            mma2 (d0, A1 + sh, B1, idesc, 1);
            mma2 (d1, A1 + sh, B2, idesc, 1);
            mma2 (d1, A2 + sh, B1, idesc, 1);
            mma2 (d2, A1 + sh, B3, idesc, 1);
            mma2 (d2, A2 + sh, B2, idesc, 1);
            if (variant >= 3)       // commit after every k-step, to a barrier nobody waits on (phases just keep completing)
                asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar2)) : "memory");
            if (variant >= 4)
                asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (variant >= 5) {     // a wait that is already satisfied: parity 1 of a fresh barrier
                unsigned done;
                asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                              : "=r"(done) : "r"(smem_u32 (&bar3)), "r"(1u) : "memory");
                if (!done) break;
            }
        }
        asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar)) : "memory");
        mbar_wait (&bar, 0);
        cycles[0] = clock64 () - t0;
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (warp == 0)
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}

static unsigned short f2bf (float f) { __nv_bfloat16 b = __float2bfloat16 (f); unsigned short u; memcpy (&u, &b, 2); return u; }
static float bf2f (unsigned short u) { unsigned v = (unsigned) u << 16; float f; memcpy (&f, &v, 4); return f; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf ("CUDA error %s at line %d\n", cudaGetErrorString (e), __LINE__); exit (1); } } while (0)

int main ()
{
    const int aRows = 136, bRows = 256, planes = 16;                    // 8 k-steps of 16
    const unsigned aPlane = aRows * 16, bPlane = bRows * 16;
    std::vector<unsigned char> aImg (planes * aPlane), bImg (planes * bPlane);
    srand (1);
    auto fill = [] (std::vector<unsigned char> &img) {
        unsigned short *p = reinterpret_cast<unsigned short *> (img.data ());
        for (size_t i = 0; i < img.size () / 2; ++i) p[i] = f2bf ((float) rand () / RAND_MAX - 0.5f);
    };
    fill (aImg); fill (bImg);
    auto elem = [] (const std::vector<unsigned char> &img, unsigned off, unsigned lbo, unsigned sbo, int row, int k) {
        const unsigned at = off + (k / 8) * lbo + (row / 8) * sbo + (row % 8) * 16 + (k % 8) * 2;
        unsigned short u; memcpy (&u, &img[at], 2); return bf2f (u);
    };

    std::vector<Op> ops;
    // columns 0..159: 8 k-steps accumulated, A shifted by (ks % 4) rows   (checks 1, 2, 4)
    for (int ks = 0; ks < 8; ++ks)
        ops.push_back ({ (unsigned) (2 * ks * aPlane + (ks % 4) * 16), (unsigned) (2 * ks * bPlane), aPlane, 128, bPlane, 128, 160, 0, (unsigned) (ks > 0) });
    // columns 160..255: fresh accumulator, then an N=48 sub-range at columns 192..239 on top (check 3)
    ops.push_back ({ 0, 160 * 16, aPlane, 128, bPlane, 128, 96, 160, 0 });
    ops.push_back ({ 2 * aPlane + 32, (unsigned) (2 * bPlane + 40 * 16), aPlane, 128, bPlane, 128, 48, 192, 1 });

    const int dCols = 256;
    std::vector<double> ref ((size_t) 128 * dCols, 0.0);
    for (const Op &o : ops)
        for (int i = 0; i < 128; ++i)
            for (unsigned j = 0; j < o.N; ++j) {
                double s = 0.0;
                for (int k = 0; k < 16; ++k)
                    s += (double) elem (aImg, o.aOff, o.aLbo, o.aSbo, i, k) * (double) elem (bImg, o.bOff, o.bLbo, o.bSbo, (int) j, k);
                double &d = ref[(size_t) i * dCols + o.dCol + j];
                d = o.acc ? d + s : s;
            }

    unsigned char *dA, *dB; Op *dOps; float *dD; long long *dCyc;
    CK (cudaMalloc (&dA, aImg.size ())); CK (cudaMalloc (&dB, bImg.size ()));
    CK (cudaMalloc (&dOps, ops.size () * sizeof (Op))); CK (cudaMalloc (&dD, 128 * dCols * 4)); CK (cudaMalloc (&dCyc, 16));
    CK (cudaMemcpy (dA, aImg.data (), aImg.size (), cudaMemcpyHostToDevice));
    CK (cudaMemcpy (dB, bImg.data (), bImg.size (), cudaMemcpyHostToDevice));
    CK (cudaMemcpy (dOps, ops.data (), ops.size () * sizeof (Op), cudaMemcpyHostToDevice));
    const size_t smemBytes = ((aImg.size () + 1023) & ~(size_t) 1023) + bImg.size () + 1024 + 8192;
    CK (cudaFuncSetAttribute (probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemBytes));

    probe<<<1, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, (int) ops.size (), dD, dCols, 0, 160, dCyc, 0, 0);
    CK (cudaDeviceSynchronize ());
    std::vector<float> got ((size_t) 128 * dCols);
    CK (cudaMemcpy (got.data (), dD, got.size () * 4, cudaMemcpyDeviceToHost));
    double worstA = 0, worstB = 0;
    int badA = 0, badB = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < dCols; ++j) {
            const double e = fabs ((double) got[(size_t) i * dCols + j] - ref[(size_t) i * dCols + j]);
            if (j < 160) { if (e > worstA) worstA = e; if (e > 1e-4) ++badA; }
            else         { if (e > worstB) worstB = e; if (e > 1e-4) ++badB; }
        }
    printf ("layout check  (8 k-steps, shifted A, N=160): max |err| %.3g, %d of %d off by > 1e-4\n", worstA, badA, 128 * 160);
    printf ("sub-range check (N=96 then N=48 at +32)     : max |err| %.3g, %d of %d off by > 1e-4\n", worstB, badB, 128 * 96);
    printf ("sample: got %.6f %.6f %.6f  ref %.6f %.6f %.6f\n", got[0], got[1], got[5 * dCols + 200], ref[0], ref[1], ref[5 * dCols + 200]);


    // accumulation behaviour: 64 and 216 k-steps into one accumulator; signed error in ulps of the result
    for (int steps : { 8, 64, 216 }) {
        std::vector<Op> chain;
        for (int ks = 0; ks < steps; ++ks)
            chain.push_back ({ (unsigned) (2 * (ks % 8) * aPlane + ((ks / 8) % 4) * 16), (unsigned) (2 * ((ks * 3) % 8) * bPlane), aPlane, 128, bPlane, 128, 160, 0, (unsigned) (ks > 0) });
        std::vector<double> r2 ((size_t) 128 * 160, 0.0), mag ((size_t) 128 * 160, 0.0);
        for (const Op &o : chain)
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 160; ++j) {
                    double s = 0.0;
                    for (int k = 0; k < 16; ++k)
                        s += (double) elem (aImg, o.aOff, o.aLbo, o.aSbo, i, k) * (double) elem (bImg, o.bOff, o.bLbo, o.bSbo, j, k);
                    r2[(size_t) i * 160 + j] += s;
                }
        Op *dChain; CK (cudaMalloc (&dChain, chain.size () * sizeof (Op)));
        CK (cudaMemcpy (dChain, chain.data (), chain.size () * sizeof (Op), cudaMemcpyHostToDevice));
        probe<<<1, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dChain, (int) chain.size (), dD, dCols, 0, 160, dCyc, 0, 0);
        CK (cudaDeviceSynchronize ());
        CK (cudaMemcpy (got.data (), dD, got.size () * 4, cudaMemcpyDeviceToHost));
        double sumSigned = 0, sumAbs = 0, worst = 0; int n = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 160; ++j) {
                const double r = r2[(size_t) i * 160 + j], g = got[(size_t) i * dCols + j];
                if (fabs (r) < 0.5) continue;
                const double ulp = ldexp (1.0, (int) floor (log2 (fabs (r))) - 23);
                const double e = (g - r) / ulp * (r > 0 ? 1.0 : -1.0);        // > 0: magnitude too large
                sumSigned += e; sumAbs += fabs (e); if (fabs (e) > worst) worst = fabs (e); ++n;
            }
        printf ("accumulate %3d k-steps: mean signed err %+.3f ulp, mean |err| %.3f ulp, max %.2f ulp (%d results with |y| >= 0.5)\n",
                steps, sumSigned / n, sumAbs / n, worst, n);
        cudaFree (dChain);
    }
    // integer-valued operands (|v| <= 127): sums stay below 2^24, the fp32 accumulation must be exact
    {
        std::vector<unsigned char> aInt (aImg.size ()), bInt (bImg.size ());
        auto fillInt = [] (std::vector<unsigned char> &img) {
            unsigned short *p = reinterpret_cast<unsigned short *> (img.data ());
            for (size_t i = 0; i < img.size () / 2; ++i) p[i] = f2bf ((float) (rand () % 255 - 127));
        };
        fillInt (aInt); fillInt (bInt);
        std::vector<Op> chain;
        for (int ks = 0; ks < 36; ++ks)
            chain.push_back ({ (unsigned) (2 * (ks % 8) * aPlane + ((ks / 8) % 4) * 16), (unsigned) (2 * ((ks * 3) % 8) * bPlane), aPlane, 128, bPlane, 128, 160, 0, (unsigned) (ks > 0) });
        std::vector<double> r2 ((size_t) 128 * 160, 0.0);
        for (const Op &o : chain)
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 160; ++j) {
                    double s = 0.0;
                    for (int k = 0; k < 16; ++k)
                        s += (double) elem (aInt, o.aOff, o.aLbo, o.aSbo, i, k) * (double) elem (bInt, o.bOff, o.bLbo, o.bSbo, j, k);
                    r2[(size_t) i * 160 + j] += s;
                }
        unsigned char *dA2, *dB2; Op *dChain;
        CK (cudaMalloc (&dA2, aInt.size ())); CK (cudaMalloc (&dB2, bInt.size ())); CK (cudaMalloc (&dChain, chain.size () * sizeof (Op)));
        CK (cudaMemcpy (dA2, aInt.data (), aInt.size (), cudaMemcpyHostToDevice));
        CK (cudaMemcpy (dB2, bInt.data (), bInt.size (), cudaMemcpyHostToDevice));
        CK (cudaMemcpy (dChain, chain.data (), chain.size () * sizeof (Op), cudaMemcpyHostToDevice));
        probe<<<1, 128, smemBytes>>> (dA2, (unsigned) aInt.size (), dB2, (unsigned) bInt.size (), dChain, (int) chain.size (), dD, dCols, 0, 160, dCyc, 0, 0);
        CK (cudaDeviceSynchronize ());
        CK (cudaMemcpy (got.data (), dD, got.size () * 4, cudaMemcpyDeviceToHost));
        int inexact = 0; double big = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 160; ++j) {
                if ((double) got[(size_t) i * dCols + j] != r2[(size_t) i * 160 + j]) ++inexact;
                if (fabs (r2[(size_t) i * 160 + j]) > big) big = fabs (r2[(size_t) i * 160 + j]);
            }
        printf ("integer operands, 36 k-steps: %d of %d results inexact (largest |sum| %.0f)\n", inexact, 128 * 160, big);
    }
    // issue rate with concurrent shared-memory store traffic from three other warps
    for (int bg : { 1, 2, 4 }) {
        const int iters = 4096;
        probe<<<1, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, 0, dD, 32, iters, 160, dCyc, bg, 0);
        CK (cudaDeviceSynchronize ());
        long long cyc[2]; CK (cudaMemcpy (cyc, dCyc, 16, cudaMemcpyDeviceToHost));
        printf ("N=160 with background STS.128 (throttle %d): %.1f cycles per MMA, stores %.1f B/clk\n", bg, (double) cyc[0] / iters, (double) cyc[1] / cyc[0]);
    }
    for (int N : { 64, 160, 256 }) {
        const int iters = 4096;
        probe<<<1, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, 0, dD, 32, iters, N, dCyc, 0, 0);
        CK (cudaDeviceSynchronize ());
        long long cyc; CK (cudaMemcpy (&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
        printf ("issue rate M=128 N=%3d K=16 bf16: %.1f cycles per MMA (%.0f MAC/clk; floor 128*N/256 = %d)\n", N, (double) cyc / iters,
                128.0 * N * 16 * iters / cyc, 128 * N / 256);
    }
    for (int pat : { 0, 1, 2 }) {
        const int iters = 4095;
        probe<<<1, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, 0, dD, 32, iters, 160, dCyc, 0, pat);
        CK (cudaDeviceSynchronize ());
        long long cyc; CK (cudaMemcpy (&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
        printf ("accumulator pattern %d (0: one D, 1: D = 0,160,160,320,320, 2: + commit every 5): %.1f cycles per MMA\n", pat, (double) cyc / iters);
    }
    CK (cudaFuncSetAttribute (probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int fmt16 : { 1 })
        for (unsigned aLbo : { 2112u })
            for (int variant : { 2, 3, 4, 5 }) {
                const int iters = 1000;
                probe2<<<1, 128, 200 * 1024>>> (variant, iters, aLbo, fmt16, dCyc);
                CK (cudaDeviceSynchronize ());
                long long cyc; CK (cudaMemcpy (&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
                printf ("lean loop %s A-plane pitch %u variant %d (3: +commit per 5, 4: +fence, 5: +satisfied try_wait): %.1f cycles per MMA\n", fmt16 ? "fp16" : "bf16", aLbo, variant, (double) cyc / (5.0 * iters));
            }
    // all SMs at once: the same loop on 148 CTAs, wall-clock rate
    {
        const int iters = 1 << 16, N = 160;
        cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
        probe<<<148, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, 0, dD, 32, iters, N, dCyc, 0, 0);
        cudaEventRecord (e0);
        probe<<<148, 128, smemBytes>>> (dA, (unsigned) aImg.size (), dB, (unsigned) bImg.size (), dOps, 0, dD, 32, iters, N, dCyc, 0, 0);
        cudaEventRecord (e1); CK (cudaDeviceSynchronize ());
        float ms; cudaEventElapsedTime (&ms, e0, e1);
        printf ("148 CTAs x %d MMAs (N=160): %.3f ms -> %.1f TMAC/s bf16 (%.2f PFLOP/s)\n", iters, ms, 148.0 * iters * 128 * N * 16 / ms / 1e9,
                2 * 148.0 * iters * 128 * N * 16 / ms / 1e12);
    }
    return 0;
}
