/*
 * art_decimate.cu -- the float <-> integer stages either side of the resampling path (reference decimator.c) on the GPU.
 *
 *   decimate   float -> 8/16/24-bit little-endian integers with optional TPDF dither and noise shaping
 *              (decimator.c:170-199).  The quantiser sits inside the shaping feedback loop and the dither generator is
 *              a serial recurrence: one thread walks one channel, channels and contexts are the parallel dimension
 *              (inputs are read eight frames ahead so that the walk is bound by its arithmetic chain, not by memory).
 *              Without dither and shaping every sample is independent: one thread per sample.
 *   floatIntegers   the lossless inverse (decimator.c:416-450): one thread per sample.
 *
 * All float arithmetic is spelled with the non-contracting intrinsics in the reference's operation order: output bytes
 * and clipped-sample counts are bit-identical to the reference (tests/test_gpu_decimator.py).
 */
#include <vector>
#include "art_kernels.cuh"
#include "art_device.h"

namespace {

__device__ __forceinline__ unsigned int lcg15 (unsigned int x) { return ((x << 4) - x) ^ 1u; }

/* tpdf_dither, decimator.c:361-373: -1 <= n < 1; type -1 / 0 / +1 = negative / no / positive intersample correlation */
__device__ __forceinline__ artsample_t tpdf (unsigned int &gen, int type)
{
    unsigned int r = lcg15 (lcg15 (gen));
    const unsigned int first = type ? (gen ^ (unsigned int) (type >> 31)) : ~r;
    r = lcg15 (lcg15 (lcg15 (r)));
    gen = r;
    return (artsample_t) (((double) ((first >> 1) + (r >> 1)) / 2147483648.0) - 1.0);
}

/* quantise one sample (decimator.c:176-198); returns the container's low `used` bytes in an int */
__device__ __forceinline__ int quantise (const ArtDecLane &L, artsample_t x, artsample_t dith, artsample_t feedback, artsample_t &code, int &clipped)
{
    const int top = (1 << (L.bits - 1)) - 1, bottom = ~top;
    code = art_sub (art_mul (x, L.scaler), feedback);
    int v = (int) floor ((double) art_add (code, dith) + 0.5);
    clipped = 0;
    return v > top ? (clipped = 1, top) : (v < bottom ? (clipped = 1, bottom) : v);
}

__device__ __forceinline__ void put_sample (const ArtDecLane &L, unsigned char *o, int v)
{
    const int shl = (24 - L.bits) % 8, bias = (L.bits <= 8) * 128;
    for (int j = 0; j < L.pad; ++j) *o++ = 0;
    v = (int) (((unsigned int) v << shl) + (unsigned int) bias);
    *o++ = (unsigned char) v;
    if (L.bits > 8) { *o++ = (unsigned char) (v >> 8); if (L.bits > 16) *o++ = (unsigned char) (v >> 16); }
}

/* one thread per channel: the serial walk */
__global__ void __launch_bounds__ (64)
art_decimate_serial_kernel (ArtDecLane *__restrict__ lanes, int numLanes)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= numLanes) return;
    ArtDecLane L = lanes[id];
    artsample_t xs[4] = { L.x[0], L.x[1], L.x[2], L.x[3] }, ys[4] = { L.y[0], L.y[1], L.y[2], L.y[3] };
    artsample_t feedback = L.feedback;
    unsigned int rng = L.rng;
    int clips = 0;
    const artsample_t *in = L.in;
    unsigned char *out = L.out;
    for (int i0 = 0; i0 < L.frames; i0 += 8) {
        artsample_t xin[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) xin[r] = i0 + r < L.frames ? __ldg (in + (size_t) (i0 + r) * L.inStride) : (artsample_t) 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (i0 + r >= L.frames) break;
            const artsample_t dith = L.dither ? tpdf (rng, L.ditherType) : (artsample_t) 0;
            artsample_t code;
            int clipped;
            /* the reference quantises first and clips afterwards; the shaper sees the UNCLIPPED value (decimator.c:183-195) */
            const int top = (1 << (L.bits - 1)) - 1, bottom = ~top;
            code = art_sub (art_mul (xin[r], L.scaler), feedback);
            int v = (int) floor ((double) art_add (code, dith) + 0.5);
            if (L.shaping) {                                  /* biquad_apply_sample, biquad.c:78-102 */
                const artsample_t e = art_sub ((artsample_t) v, code);
                artsample_t sum = art_mul (e, L.a[0]);
                for (int d = L.order; d >= 1; --d)
                    sum = art_add (sum, art_sub (art_mul (xs[d - 1], L.a[d]), art_mul (L.b[d], ys[d - 1])));
                xs[3] = xs[2]; xs[2] = xs[1]; xs[1] = xs[0]; xs[0] = e;
                ys[3] = ys[2]; ys[2] = ys[1]; ys[1] = ys[0]; ys[0] = sum;
                feedback = sum;
            }
            clipped = v > top || v < bottom;
            v = v > top ? top : (v < bottom ? bottom : v);
            clips += clipped;
            put_sample (L, out + (size_t) (i0 + r) * L.outStride, v);
        }
    }
    ArtDecLane &W = lanes[id];
    W.rng = rng; W.feedback = feedback; W.clips = clips;
#pragma unroll
    for (int k = 0; k < 4; ++k) { W.x[k] = xs[k]; W.y[k] = ys[k]; }
}

/* no dither, no shaping: samples are independent -- one thread per (lane, frame) */
__global__ void __launch_bounds__ (256)
art_decimate_parallel_kernel (ArtDecLane *__restrict__ lanes, int numLanes, int maxFrames)
{
    const int lane = blockIdx.y;
    const ArtDecLane &L = lanes[lane];
    int clips = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.frames; i += gridDim.x * blockDim.x) {
        artsample_t code;
        int clipped;
        const int v = quantise (L, __ldg (L.in + (size_t) i * L.inStride), (artsample_t) 0, L.feedback, code, clipped);
        clips += clipped;
        put_sample (L, L.out + (size_t) i * L.outStride, v);
    }
    for (int sh = 16; sh >= 1; sh >>= 1) clips += __shfl_xor_sync (0xffffffffu, clips, sh);
    if ((threadIdx.x & 31) == 0 && clips) atomicAdd (&lanes[lane].clips, clips);
}

/* floatIntegersLE, decimator.c:416-450 */
__global__ void __launch_bounds__ (256)
art_float_integers_kernel (const unsigned char *__restrict__ in, artsample_t gain, int bits, int bytes, int stride, artsample_t *__restrict__ out, int count)
{
    const int used = (bits + 7) / 8;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long) gridDim.x * blockDim.x) {
        const unsigned char *p = in + (size_t) i * stride * bytes + (bytes - used);
        int v;
        if (bits <= 8) v = (int) p[0] - 128;
        else if (bits <= 16) v = (short) (p[0] | (p[1] << 8));
        else v = p[0] | (p[1] << 8) | ((int) (signed char) p[2] << 16);
        out[i] = art_mul ((artsample_t) v, gain);
    }
}

}   // namespace

extern "C" int artDecimateRun (ArtDecLane *lanes, int numLanes, int numContexts, int channelsHint, int onDevice, void *streamPtr)
{
    ART_GUARD_BEGIN
    (void) numContexts; (void) channelsHint;
    int count = 0;
    if (cudaGetDeviceCount (&count) != cudaSuccess || count == 0)
        artRaise ("the decimator needs a CUDA device; this library has no CPU path");
    cudaStream_t stream = (cudaStream_t) streamPtr;
    std::vector<ArtDecLane> dl (lanes, lanes + numLanes);
    unsigned char *scratch = nullptr;
    struct Piece { const void *hostIn; void *hostOut; size_t inBytes, outBytes, inAt, outAt; };
    std::vector<Piece> pieces;
    int maxFrames = 0;
    bool serial = false;
    for (int i = 0; i < numLanes; ++i) {
        dl[i].clips = 0;
        if (dl[i].frames > maxFrames) maxFrames = dl[i].frames;
        serial |= dl[i].dither || dl[i].shaping;
    }
    if (maxFrames == 0) {
        for (int i = 0; i < numLanes; ++i) lanes[i].clips = 0;
        return 0;
    }
    if (!onDevice) {
        // interleaved contexts travel as one block (their lanes are consecutive and share it), planar channels one by one
        size_t at = 0;
        std::vector<int> pieceOf (numLanes, -1);
        for (int i = 0; i < numLanes; ) {
            const ArtDecLane &L = lanes[i];
            int e = i + 1;
            if (L.inStride > 1)
                while (e < numLanes && lanes[e].context == L.context) ++e;
            Piece p;
            p.hostIn = L.in; p.hostOut = L.out;
            p.inBytes = (size_t) L.frames * L.inStride * sizeof (artsample_t);
            p.outBytes = (size_t) L.frames * L.outStride;
            p.inAt = at; at += (p.inBytes + 255) & ~(size_t) 255;
            p.outAt = at; at += (p.outBytes + 255) & ~(size_t) 255;
            for (int q = i; q < e; ++q) pieceOf[q] = (int) pieces.size ();
            pieces.push_back (p);
            i = e;
        }
        ART_CUDA_CHECK (cudaMallocAsync (&scratch, at ? at : 256, stream));
        for (const Piece &p : pieces)
            if (p.inBytes)
                ART_CUDA_CHECK (cudaMemcpyAsync (scratch + p.inAt, p.hostIn, p.inBytes, cudaMemcpyHostToDevice, stream));
        for (int i = 0; i < numLanes; ++i) {
            const Piece &p = pieces[pieceOf[i]];
            dl[i].in = reinterpret_cast<const artsample_t *> (scratch + p.inAt) + (lanes[i].in - reinterpret_cast<const artsample_t *> (p.hostIn));
            dl[i].out = scratch + p.outAt + (lanes[i].out - reinterpret_cast<unsigned char *> (p.hostOut));
        }
    }
    ArtDecLane *d_lanes = nullptr;
    ART_CUDA_CHECK (cudaMallocAsync (&d_lanes, sizeof (ArtDecLane) * numLanes, stream));
    ART_CUDA_CHECK (cudaMemcpyAsync (d_lanes, dl.data (), sizeof (ArtDecLane) * numLanes, cudaMemcpyHostToDevice, stream));
    if (serial || numLanes > 65535)
        art_decimate_serial_kernel<<<(numLanes + 63) / 64, 64, 0, stream>>> (d_lanes, numLanes);
    else {
        int bx = (maxFrames + 255) / 256;
        if (bx > 1024) bx = 1024;
        art_decimate_parallel_kernel<<<dim3 (bx, numLanes), 256, 0, stream>>> (d_lanes, numLanes, maxFrames);
    }
    ART_CUDA_CHECK (cudaGetLastError ());
    ++g_artLaunches;
    ART_CUDA_CHECK (cudaMemcpyAsync (dl.data (), d_lanes, sizeof (ArtDecLane) * numLanes, cudaMemcpyDeviceToHost, stream));
    for (const Piece &p : pieces)
        if (p.outBytes)
            ART_CUDA_CHECK (cudaMemcpyAsync (p.hostOut, scratch + p.outAt, p.outBytes, cudaMemcpyDeviceToHost, stream));
    ART_CUDA_CHECK (cudaStreamSynchronize (stream));
    ART_CUDA_CHECK (cudaFreeAsync (d_lanes, stream));
    if (scratch) ART_CUDA_CHECK (cudaFreeAsync (scratch, stream));
    for (int i = 0; i < numLanes; ++i) {
        lanes[i].rng = dl[i].rng; lanes[i].feedback = dl[i].feedback; lanes[i].clips = dl[i].clips;
        for (int k = 0; k < 4; ++k) { lanes[i].x[k] = dl[i].x[k]; lanes[i].y[k] = dl[i].y[k]; }
    }
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artFloatIntegersRun (const unsigned char *input, double gain, int bits, int bytes, int stride, artsample_t *output, int count,
                                    int onDevice, void *streamPtr)
{
    ART_GUARD_BEGIN
    if (count <= 0 || bits > 24) return 0;              // (the reference does nothing above 24 bits either)
    cudaStream_t stream = (cudaStream_t) streamPtr;
    const artsample_t g = (artsample_t) (bits <= 8 ? gain / 128.0 : (bits <= 16 ? gain / 32768.0 : gain / 8388608.0));
    const size_t inBytes = (size_t) count * stride * bytes;     // the last sample's trailing channels are never read, but belong to the block
    const unsigned char *d_in = input;
    artsample_t *d_out = output;
    unsigned char *scratch = nullptr;
    if (!onDevice) {
        const size_t inRound = (inBytes + 255) & ~(size_t) 255;
        ART_CUDA_CHECK (cudaMallocAsync (&scratch, inRound + (size_t) count * sizeof (artsample_t), stream));
        // only (count - 1) * stride * bytes + bytes bytes are guaranteed to exist behind `input`
        ART_CUDA_CHECK (cudaMemcpyAsync (scratch, input, (size_t) (count - 1) * stride * bytes + bytes, cudaMemcpyHostToDevice, stream));
        d_in = scratch;
        d_out = reinterpret_cast<artsample_t *> (scratch + inRound);
    }
    int blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    art_float_integers_kernel<<<blocks, 256, 0, stream>>> (d_in, g, bits, bytes, stride, d_out, count);
    ART_CUDA_CHECK (cudaGetLastError ());
    ++g_artLaunches;
    if (!onDevice) {
        ART_CUDA_CHECK (cudaMemcpyAsync (output, d_out, (size_t) count * sizeof (artsample_t), cudaMemcpyDeviceToHost, stream));
        ART_CUDA_CHECK (cudaStreamSynchronize (stream));
        ART_CUDA_CHECK (cudaFreeAsync (scratch, stream));
    }
    return 0;
    ART_GUARD_END (-1)
}
