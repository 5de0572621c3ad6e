/*
 * art_device.cu -- device-side context of a resampler: history, filter bank, staging,
 * job construction and kernel launches.  Implements art_device.h.
 *
 * What it replaces in the reference: the per-channel ring buffers (resampler.c:171-174),
 * their compaction (:497-503, :614-620), the zero post-fill of a flush (:663-685) and the
 * per-channel worker threads of workers.c (channels are simply a grid dimension here).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdarg>
#include <time.h>
#include <mutex>
#include <sched.h>
#include <vector>
#include "art_kernels.cuh"
#include "art_device.h"

std::atomic<unsigned long long> g_artLaunches { 0 };
static std::atomic<unsigned long long> g_pathLaunches[3];       // [0] generic kernel, [1] periodic FFMA kernels, [2] tensor-core kernel

/* ---- error model (art_kernels.cuh): message of the last failure on this thread ------------------------------- */
static thread_local char g_lastError[512];
static thread_local bool g_hasError = false;

void artNote (const char *what)
{
    snprintf (g_lastError, sizeof g_lastError, "%s", what);
    g_hasError = true;
    fprintf (stderr, "libresampler_b200: %s\n", g_lastError);
}

void artRaiseCuda (cudaError_t code, const char *file, int line, const char *expr)
{
    char msg[512];
    const char *base = strrchr (file, '/');
    snprintf (msg, sizeof msg, "CUDA error %s at %s:%d (%s)", cudaGetErrorString (code), base ? base + 1 : file, line, expr);
    artNote (msg);
    throw ArtError { (int) code };
}

void artRaise (const char *fmt, ...)
{
    char msg[512];
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (msg, sizeof msg, fmt, ap);
    va_end (ap);
    artNote (msg);
    throw ArtError { -1 };
}

extern "C" const char *artDevLastError (int clear)
{
    if (!g_hasError) return nullptr;
    if (clear) g_hasError = false;
    return g_lastError;
}

extern "C" void artDevPathCounts (unsigned long long *generic, unsigned long long *periodic)
{
    if (generic) *generic = g_pathLaunches[0];
    if (periodic) *periodic = g_pathLaunches[1];
}

extern "C" unsigned long long artDevTensorLaunches (void) { return g_pathLaunches[2]; }

int g_artTensorMode = -1;          // -1: read ART_B200_UMMA on first use; 0 off, 1 when the launch is large enough, 2 whenever eligible, 3 as 1 plus non-interpolating contexts
extern "C" void artDevSetTensorMode (int mode) { g_artTensorMode = mode < 0 ? 0 : (mode > 3 ? 3 : mode); }

extern "C" void artDevSetTensorDigits (int digits) { g_artTensorDigits = digits == 2 ? 2 : 3; }

extern "C" unsigned long long artDevLaunchCount (void) { return g_artLaunches; }

/* ---- optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching
 * stream right around the convolution kernels, collected after the timed region ------------------ */
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profEvents;
static std::mutex g_profMutex;

void artProfileBegin (cudaStream_t stream, void **token)
{
    *token = nullptr;
    if (!g_profile) return;
    cudaEvent_t a, b;
    ART_CUDA_CHECK (cudaEventCreate (&a));
    ART_CUDA_CHECK (cudaEventCreate (&b));
    ART_CUDA_CHECK (cudaEventRecord (a, stream));
    std::lock_guard<std::mutex> lock (g_profMutex);
    g_profEvents.emplace_back (a, b);
    *token = b;
}

void artProfileEnd (cudaStream_t stream, void *token)
{
    if (token) ART_CUDA_CHECK (cudaEventRecord ((cudaEvent_t) token, stream));
}

extern "C" void artDevProfileEnable (int on) { g_profile = on != 0; }

extern "C" unsigned long long artDevProfileCollect (double *totalMs)
{
    if (totalMs) *totalMs = 0.0;
    ART_GUARD_BEGIN
    std::lock_guard<std::mutex> lock (g_profMutex);
    double sum = 0.0;
    for (auto &p : g_profEvents) {
        float ms = 0.0f;
        ART_CUDA_CHECK (cudaEventSynchronize (p.second));
        ART_CUDA_CHECK (cudaEventElapsedTime (&ms, p.first, p.second));
        sum += ms;
        cudaEventDestroy (p.first);
        cudaEventDestroy (p.second);
    }
    const unsigned long long n = g_profEvents.size ();
    g_profEvents.clear ();
    if (totalMs) *totalMs = sum;
    return n;
    ART_GUARD_END (0)
}

/* ---- filter banks are immutable and identical for equal (T, F, coefficients): share them ---- */
struct ArtBank {
    int device, T, Tp, F, refs;
    float absSum;
    unsigned long long hash;
    artsample_t *d_rows;
    std::vector<artsample_t> packed;
};

static std::mutex g_bankMutex;
static std::vector<ArtBank *> g_banks;

static ArtBank *bank_acquire (int device, int T, int F, const artsample_t *const *rows)
{
    const int Tp = (T + 31) & ~31;
    std::vector<artsample_t> packed ((size_t) (F + 2) * Tp, (artsample_t) 0);
    unsigned long long h = 1469598103934665603ULL;
    for (int r = 0; r <= F; ++r) {
        memcpy (&packed[(size_t) r * Tp], rows[r], sizeof (artsample_t) * T);
        const unsigned char *b = reinterpret_cast<const unsigned char *> (rows[r]);
        for (size_t i = 0; i < sizeof (artsample_t) * T; ++i) { h ^= b[i]; h *= 1099511628211ULL; }
    }
    std::lock_guard<std::mutex> lock (g_bankMutex);
    for (ArtBank *b : g_banks)
        if (b->device == device && b->T == T && b->F == F && b->hash == h && b->packed == packed) {
            b->refs++;
            return b;
        }
    ArtBank *b = new ArtBank;
    b->device = device; b->T = T; b->Tp = Tp; b->F = F; b->refs = 1; b->hash = h;
    b->packed.swap (packed);
    b->absSum = 0.0f;
    for (int r = 0; r <= F; ++r) {
        double s = 0.0;
        for (int t = 0; t < T; ++t) s += fabs ((double) b->packed[(size_t) r * Tp + t]);
        if (s > b->absSum) b->absSum = (float) s;
    }
    if (cudaMalloc (&b->d_rows, b->packed.size () * sizeof (artsample_t)) != cudaSuccess ||
        cudaMemcpy (b->d_rows, b->packed.data (), b->packed.size () * sizeof (artsample_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        fprintf (stderr, "libresampler_b200: cannot place the filter bank on the GPU: %s\n",
                 cudaGetErrorString (cudaGetLastError ()));
        delete b;
        return nullptr;
    }
    g_banks.push_back (b);
    return b;
}

static void bank_release (ArtBank *bank)
{
    std::lock_guard<std::mutex> lock (g_bankMutex);
    if (--bank->refs > 0)
        return;
    for (size_t i = 0; i < g_banks.size (); ++i)
        if (g_banks[i] == bank) { g_banks.erase (g_banks.begin () + i); break; }
    cudaFree (bank->d_rows);
    delete bank;
}

/* ---- context ---------------------------------------------------------------------------------- */
struct ArtDev {
    int device, smCount;
    int C, T, F, mode;
    ArtBank *bank;
    cudaStream_t stream;
    cudaStream_t lastStream;        // where the most recent device-pointer call was enqueued (reset has to wait for it)
    artsample_t *hist[2];
    int cur;
    artsample_t *d_in, *d_out;
    size_t inCap, outCap;           // floats
    artsample_t *d_stage;                 // flush block of the endpoint extrapolation (device-pointer calls)
    size_t stageCap;
    ArtClass klass;
    // host-pointer path: copies in, kernels and copies out run on three streams so that PCIe traffic in
    // both directions overlaps the convolution (created on first use)
    cudaStream_t sIn, sOut;
    std::vector<cudaEvent_t> events;
    cudaEvent_t doneEvent;          // blocking-sync event: large host calls sleep on it instead of spinning on the stream
    // phase tables of single-job periodic launches: reused from call to call (calls on one context are
    // stream-ordered), so the latency path has no allocation in it
    void *tableBuf;
    size_t tableCap;
};

static void host_pipe_init (ArtDev *dev, size_t events)
{
    if (!dev->sIn) {
        ART_CUDA_CHECK (cudaStreamCreateWithFlags (&dev->sIn, cudaStreamNonBlocking));
        ART_CUDA_CHECK (cudaStreamCreateWithFlags (&dev->sOut, cudaStreamNonBlocking));
    }
    while (dev->events.size () < events) {
        cudaEvent_t e;
        ART_CUDA_CHECK (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
        dev->events.push_back (e);
    }
}

/* Wait for a host call's last transfer.  cudaStreamSynchronize spins inside the driver (the runtime's default schedule): the
 * lowest latency, and what keeps PCIe busy when a process has several calls in flight from several threads -- a sleeping wait
 * costs a wake-up per call (one GPU, 16 threads, 2 MB calls: 9.1 Gsamples/s end to end against 11.3-11.4 spinning; polling
 * cudaEventQuery from every thread is no better, 8.5: the queries contend with the other threads' submissions).  Spinning needs a
 * core per waiting thread, so the default (mode 2) spins while the process has no more waiting threads than CPUs it may run on
 * and sleeps on a blocking-sync event otherwise; eight ranks x 4 threads on a 32-CPU host measured the same under either
 * (17.6 / 17.7 Gsamples/s, profiles/r02_e2e_scale_n8.txt: the host's memory system is the limit there). */
static int g_waitMode = -1;         // ART_B200_WAIT: "spin" 0, "block" 1 (sleep on a blocking-sync event), default 2 (spin while there are cores for it)
static std::atomic<int> g_waiters { 0 };
static int g_cpus = 0;

static void wait_for (ArtDev *dev, cudaStream_t stream, size_t bytesMoved)
{
    if (g_waitMode < 0) {
        const char *e = getenv ("ART_B200_WAIT");
        cpu_set_t set;
        g_cpus = sched_getaffinity (0, sizeof set, &set) == 0 ? CPU_COUNT (&set) : 1;
        g_waitMode = !e ? 2 : (!strcmp (e, "spin") ? 0 : (!strcmp (e, "block") ? 1 : 2));
    }
    bool spin = bytesMoved < (1u << 20) || g_waitMode == 0;
    if (!spin && g_waitMode == 2) {
        const int now = ++g_waiters;
        spin = now <= g_cpus;
        if (spin) {
            const cudaError_t e_ = cudaStreamSynchronize (stream);
            --g_waiters;
            ART_CUDA_CHECK (e_);
            return;
        }
        --g_waiters;
    }
    if (spin) {
        ART_CUDA_CHECK (cudaStreamSynchronize (stream));
        return;
    }
    if (!dev->doneEvent)
        ART_CUDA_CHECK (cudaEventCreateWithFlags (&dev->doneEvent, cudaEventBlockingSync | cudaEventDisableTiming));
    ART_CUDA_CHECK (cudaEventRecord (dev->doneEvent, stream));
    ART_CUDA_CHECK (cudaEventSynchronize (dev->doneEvent));
}

static void use_device (const ArtDev *dev)
{
    int now = -1;
    if (cudaGetDevice (&now) != cudaSuccess || now != dev->device)
        ART_CUDA_CHECK (cudaSetDevice (dev->device));
}

extern "C" ArtDev *artDevCreate (int channels, int taps, int lead, int filters, int mode, const artsample_t *const *rows)
{
    ART_GUARD_BEGIN
    int device = 0, count = 0;
    cudaError_t e = cudaGetDeviceCount (&count);
    if (e != cudaSuccess || count == 0) {
        fprintf (stderr, "libresampler_b200: no usable CUDA device (%s); this library has no CPU path\n",
                 e != cudaSuccess ? cudaGetErrorString (e) : "device count is 0");
        return nullptr;
    }
    if (cudaGetDevice (&device) != cudaSuccess) {
        fprintf (stderr, "libresampler_b200: cudaGetDevice failed: %s\n", cudaGetErrorString (cudaGetLastError ()));
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties (&prop, device) != cudaSuccess || prop.major < 10) {
        fprintf (stderr, "libresampler_b200: device %d is not a Blackwell (sm_100a) GPU; kernels are built for sm_100a only\n", device);
        return nullptr;
    }

    {
        // the per-launch scratch (job arrays, phase tables) comes from the stream-ordered pool; keep what
        // it has instead of returning it to the driver at every synchronisation (default threshold is 0)
        static bool poolTuned[64] = { false };
        if (device < 64 && !poolTuned[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool (&pool, device) == cudaSuccess) {
                unsigned long long keep = ~0ULL;
                cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            poolTuned[device] = true;
        }
    }

    ArtDev *dev = new ArtDev ();
    dev->device = device;
    dev->smCount = prop.multiProcessorCount;
    dev->C = channels; dev->T = taps; dev->F = filters; dev->mode = mode;
    dev->bank = bank_acquire (device, taps, filters, rows);
    if (!dev->bank) { delete dev; return nullptr; }
    ART_CUDA_CHECK (cudaStreamCreateWithFlags (&dev->stream, cudaStreamNonBlocking));
    const size_t histBytes = sizeof (artsample_t) * (size_t) (channels > 0 ? channels : 1) * taps;
    for (int i = 0; i < 2; ++i) {
        ART_CUDA_CHECK (cudaMalloc (&dev->hist[i], histBytes));
        ART_CUDA_CHECK (cudaMemset (dev->hist[i], 0, histBytes));
    }
    dev->cur = 0;
    dev->lastStream = nullptr;
    dev->d_in = dev->d_out = nullptr;
    dev->inCap = dev->outCap = 0;
    dev->d_stage = nullptr;
    dev->stageCap = 0;
    dev->sIn = dev->sOut = nullptr;
    dev->doneEvent = nullptr;
    dev->tableBuf = nullptr;
    dev->tableCap = 0;

    ArtClass &k = dev->klass;
    memset (&k, 0, sizeof k);
    k.bank = dev->bank->d_rows;
    k.T = taps; k.Tp = dev->bank->Tp; k.F = filters; k.C = channels; k.mode = mode;
    k.Tref = taps - lead; k.lead = lead;
    k.absSum = dev->bank->absSum;
    k.sort = getenv ("ART_B200_NOSORT") ? 0 : 1;
    return dev;
    ART_GUARD_END (nullptr)
}

extern "C" void artDevDestroy (ArtDev *dev)
{
    if (!dev) return;
    use_device (dev);
    cudaStreamSynchronize (dev->stream);
    cudaFree (dev->hist[0]);
    cudaFree (dev->hist[1]);
    cudaFree (dev->d_in);
    cudaFree (dev->d_out);
    cudaFree (dev->d_stage);
    cudaFree (dev->tableBuf);
    cudaStreamDestroy (dev->stream);
    if (dev->sIn) { cudaStreamDestroy (dev->sIn); cudaStreamDestroy (dev->sOut); }
    for (cudaEvent_t e : dev->events) cudaEventDestroy (e);
    if (dev->doneEvent) cudaEventDestroy (dev->doneEvent);
    bank_release (dev->bank);
    delete dev;
}

extern "C" int artDevReset (ArtDev *dev)
{
    ART_GUARD_BEGIN
    use_device (dev);
    // device-pointer calls run on the caller's stream and write the history: wait for the last one before clearing it
    // (a stream the caller has destroyed since has nothing pending: its error is dropped)
    if (dev->lastStream && dev->lastStream != dev->stream && cudaStreamSynchronize (dev->lastStream) != cudaSuccess)
        (void) cudaGetLastError ();
    dev->lastStream = nullptr;
    ART_CUDA_CHECK (cudaMemsetAsync (dev->hist[dev->cur], 0, sizeof (artsample_t) * (size_t) dev->C * dev->T, dev->stream));
    ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevDeviceIndex (const ArtDev *dev) { return dev->device; }

extern "C" int artDevSelect (int device)
{
    cudaError_t e = cudaSetDevice (device);
    if (e != cudaSuccess)
        fprintf (stderr, "libresampler_b200: cudaSetDevice(%d): %s\n", device, cudaGetErrorString (e));
    return e == cudaSuccess ? 0 : (int) e;
}

extern "C" int artDevCount (void)
{
    int n = 0;
    return cudaGetDeviceCount (&n) == cudaSuccess ? n : 0;
}

extern "C" int artDevSynchronize (ArtDev *dev)
{
    ART_GUARD_BEGIN
    use_device (dev);
    ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevGetHistory (ArtDev *dev, artsample_t *hostPlanar)
{
    ART_GUARD_BEGIN
    use_device (dev);
    ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
    ART_CUDA_CHECK (cudaMemcpy (hostPlanar, dev->hist[dev->cur], sizeof (artsample_t) * (size_t) dev->C * dev->T, cudaMemcpyDeviceToHost));
    return 0;
    ART_GUARD_END (-1)
}

/* ---- endpoint extrapolation support (art_context.c): small synchronous transfers at a stream's start and end ---- */
extern "C" int artDevGetHistoryOn (ArtDev *dev, artsample_t *hostPlanar, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    ART_CUDA_CHECK (cudaMemcpyAsync (hostPlanar, dev->hist[dev->cur], sizeof (artsample_t) * (size_t) dev->C * dev->T, cudaMemcpyDeviceToHost, st));
    ART_CUDA_CHECK (cudaStreamSynchronize (st));
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevPatchHistory (ArtDev *dev, int channel, int first, int count, const artsample_t *values, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    // pageable source: staged by the runtime before the call returns
    ART_CUDA_CHECK (cudaMemcpyAsync (dev->hist[dev->cur] + (size_t) channel * dev->T + first, values, sizeof (artsample_t) * (size_t) count,
                                     cudaMemcpyHostToDevice, st));
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevFetch (ArtDev *dev, const artsample_t *d_src, size_t floats, artsample_t *host, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    ART_CUDA_CHECK (cudaMemcpyAsync (host, d_src, sizeof (artsample_t) * floats, cudaMemcpyDeviceToHost, st));
    ART_CUDA_CHECK (cudaStreamSynchronize (st));
    return 0;
    ART_GUARD_END (-1)
}

extern "C" artsample_t *artDevStage (ArtDev *dev, const artsample_t *host, size_t floats, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    if (floats > dev->stageCap) {
        ART_CUDA_CHECK (cudaStreamSynchronize (st));
        cudaFree (dev->d_stage);
        dev->stageCap = floats + 1024;
        ART_CUDA_CHECK (cudaMalloc (&dev->d_stage, dev->stageCap * sizeof (artsample_t)));
    }
    ART_CUDA_CHECK (cudaMemcpyAsync (dev->d_stage, host, sizeof (artsample_t) * floats, cudaMemcpyHostToDevice, st));
    return dev->d_stage;
    ART_GUARD_END (nullptr)
}

extern "C" int artDevSetHistory (ArtDev *dev, const artsample_t *hostPlanar)
{
    ART_GUARD_BEGIN
    use_device (dev);
    ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
    ART_CUDA_CHECK (cudaMemcpy (dev->hist[dev->cur], hostPlanar, sizeof (artsample_t) * (size_t) dev->C * dev->T, cudaMemcpyHostToDevice));
    return 0;
    ART_GUARD_END (-1)
}

/* ---- job construction ----------------------------------------------------------------------------- */

static void fill_job (ArtDev *dev, const ArtCallPlan &p, ArtJob &j)
{
    memset (&j, 0, sizeof j);
    j.P = p.st.P;
    j.ratio = p.st.ratio;
    j.I = p.st.I;
    j.origin = p.st.I - p.pre;
    j.outputs = p.outputs;
    j.inValid = p.inValid;
    j.prevAvail = 0;
    j.consumed = p.consumed;
    j.hist = dev->hist[dev->cur];
    j.histOut = p.consumed ? dev->hist[dev->cur ^ 1] : nullptr;
}

static void finish_job (ArtDev *dev, const ArtCallPlan &p)
{
    if (p.consumed)
        dev->cur ^= 1;
}

/* ---- kernel choice and launch ------------------------------------------------------------------------
 * One "launch" covers a list of jobs of one configuration: a single call, the contexts of a batch, or
 * the blocks of an ASRC sequence.  Rational ratios with a small numerator go to the periodic kernel
 * (every job of the launch must share the ratio); everything else to the generic kernel.  Long calls
 * are cut into segments for the periodic kernel (see artPeriodicSegmentOutputs). */
struct ArtLaunchPlan {
    ArtClass k;
    bool periodic;              // one of the two rational-ratio kernels
    bool umma;                  // ... the tensor-core one (art_sinc_umma.cu)
    ArtUmma um;
    int smCount;
    ArtPeriodic per;
    int CV;
    ArtLaunchGeom g;
    unsigned int segLen;
};

static bool g_forceGeneric = false, g_envRead = false;

static void plan_launch (ArtDev *lead, double minRatio, double maxRatio, bool oneRatio, unsigned int maxOut, unsigned long long totalOut,
                         bool allowPeriodic, ArtLaunchPlan &lp)
{
    if (!g_envRead) {
        g_forceGeneric = getenv ("ART_B200_GENERIC") != nullptr;     // debugging / A-B measurements
        g_envRead = true;
    }
    lp.k = lead->klass;
    lp.periodic = false;
    lp.umma = false;
    lp.smCount = lead->smCount;
    lp.segLen = 0;
    if (!maxOut)
        return;
    if (allowPeriodic && oneRatio && !g_forceGeneric &&
        artPlanUmma (lp.k, minRatio, maxOut, totalOut, lead->smCount, lp.um)) {
        lp.periodic = lp.umma = true;
        lp.per.L = lp.um.L; lp.per.M = lp.um.M;
        lp.segLen = artPeriodicSegmentOutputs (lp.per, minRatio);
        return;
    }
    const unsigned int total32 = (unsigned int) (totalOut > 0xffffffffULL ? 0xffffffffULL : totalOut);
    if (allowPeriodic && oneRatio && !g_forceGeneric &&
        artPlanPeriodic (lp.k, minRatio, maxOut, totalOut, lead->smCount, lp.per, lp.CV)) {
        lp.periodic = true;
        lp.segLen = artPeriodicSegmentOutputs (lp.per, minRatio);
        return;
    }
    // asynchronous sample-rate conversion: ratios within a few hundred ppm of 1.  The read position then advances by almost exactly
    // one sample per output, the filter-row pair changes every 1 / (|1/r - 1| * F) outputs, and runs of consecutive outputs share it
    {
        const double dLo = fabs (1.0 / minRatio - 1.0), dHi = fabs (1.0 / maxRatio - 1.0), d = dLo > dHi ? dLo : dHi;
        lp.k.unity = !ART_WIDE && (lp.k.mode & ART_MODE_INTERP) && !(lp.k.mode & ART_MODE_PRECISE) && d * lp.k.F <= 0.125 && !getenv ("ART_B200_NOUNITY");
    }
    artPlanGenericGeometry (lp.k, minRatio, total32, lead->smCount, lp.g);
}

/* append the job (cut into segments when the periodic kernel needs that); returns CTAs/tiles added */
static int append_job (const ArtLaunchPlan &lp, const ArtJob &base, std::vector<ArtJob> &jobs, int firstCta)
{
    int ctas = 0;
    if (!base.outputs) {
        if (base.histOut) {                 // nothing to produce, but the history still has to move on
            ArtJob j = base;
            j.tile0 = firstCta;
            jobs.push_back (j);
        }
        return 0;
    }
    if (!lp.periodic) {
        ArtJob j = base;
        j.tile0 = firstCta;
        jobs.push_back (j);
        return (int) ((base.outputs + lp.k.NB - 1) / lp.k.NB);
    }
    for (unsigned int at = 0; at < base.outputs; at += lp.segLen) {
        ArtJob j = base;
        j.nStart = base.nStart + at;
        j.outputs = base.outputs - at < lp.segLen ? base.outputs - at : lp.segLen;
        j.tile0 = firstCta + ctas;
        if (at) j.histOut = nullptr;        // one history update per call
        jobs.push_back (j);
        ctas += lp.umma ? artUmmaTiles (lp.um, lp.k.C, j.outputs)
                        : artPeriodicCtas (lp.per, j.outputs);
    }
    return ctas;
}


/* ---- many-channel interleaved blocks on the tensor-core kernel -------------------------------------------
 * Its tiles are single channels; read straight from an interleaved block of C >= 8 channels every 32-byte sector would
 * carry one useful sample (and 8 tiles would fetch the same sector), and the epilogue would store 4 bytes per sector.
 * Such launches go through planar scratch in HBM instead: one coalesced transposition in, one out (+16 bytes of
 * traffic per sample on a path that sits far below the HBM roofline). */
struct ArtXpose { const artsample_t *src; artsample_t *dst; long long frames, pitch; };

__global__ void __launch_bounds__ (256)
art_deinterleave_kernel (const ArtXpose *__restrict__ g, int C, int groups)            // dst[c * pitch + f] = src[f * C + c]
{
    __shared__ artsample_t t[32][33];
    for (int z = blockIdx.z; z < groups; z += gridDim.z) {
        const ArtXpose x = g[z];
        const long long f0 = (long long) blockIdx.x * 32;
        const int c0 = blockIdx.y * 32;
        if (f0 >= x.frames) continue;
        for (int r = threadIdx.y; r < 32; r += 8) {
            const long long f = f0 + r;
            const int c = c0 + threadIdx.x;
            t[r][threadIdx.x] = (f < x.frames && c < C) ? __ldg (x.src + f * C + c) : 0.0f;
        }
        __syncthreads ();
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int c = c0 + r;
            const long long f = f0 + threadIdx.x;
            if (c < C && f < x.frames) x.dst[(long long) c * x.pitch + f] = t[threadIdx.x][r];
        }
        __syncthreads ();
    }
}

__global__ void __launch_bounds__ (256)
art_interleave_kernel (const ArtXpose *__restrict__ g, int C, int groups)              // dst[f * C + c] = src[c * pitch + f]
{
    __shared__ artsample_t t[32][33];
    for (int z = blockIdx.z; z < groups; z += gridDim.z) {
        const ArtXpose x = g[z];
        const long long f0 = (long long) blockIdx.x * 32;
        const int c0 = blockIdx.y * 32;
        if (f0 >= x.frames) continue;
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int c = c0 + r;
            const long long f = f0 + threadIdx.x;
            t[r][threadIdx.x] = (c < C && f < x.frames) ? __ldg (x.src + (long long) c * x.pitch + f) : 0.0f;
        }
        __syncthreads ();
        for (int r = threadIdx.y; r < 32; r += 8) {
            const long long f = f0 + r;
            const int c = c0 + threadIdx.x;
            if (f < x.frames && c < C) x.dst[f * C + c] = t[threadIdx.x][r];
        }
        __syncthreads ();
    }
}

/* region index of the first input frame that output n of a job reads (the start of its window) */
static long long first_frame_read (const ArtJob &j, const ArtClass &k, unsigned int n)
{
    ArtLoopState st;
    st.P = j.P; st.ratio = j.ratio; st.I = j.I; st.T = k.Tref;
    int w;
    const double pos = art_output_pos (&st, n, &w);
    return (long long) floor (pos) - (k.Tref / 2 + k.lead) + 1 + (long long) w * 15LL * k.Tref - j.origin;
}

/* Job lists travel through a small ring of pinned buffers per GPU.  A pageable source larger than the runtime's staging
 * threshold (64 KB: ~400 jobs) makes cudaMemcpyAsync wait until the stream has reached the copy, i.e. for the previous
 * launch -- an ASRC sequence of 1024 blocks per launch then alternates host and GPU work instead of overlapping them. */
struct ArtJobRing { char *host; size_t slotCap; cudaEvent_t done[8]; bool used[8]; unsigned int next; };
static ArtJobRing g_jobRing[16];
static std::mutex g_jobRingMutex;

static void upload_jobs (void *d_dst, const void *src, size_t bytes, cudaStream_t stream)
{
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    std::lock_guard<std::mutex> lock (g_jobRingMutex);
    ArtJobRing &ring = g_jobRing[device & 15];
    if (bytes > ring.slotCap) {                                   // one pinned arena of eight slots, grown rarely
        for (int i = 0; i < 8; ++i)
            if (ring.used[i]) { ART_CUDA_CHECK (cudaEventSynchronize (ring.done[i])); ring.used[i] = false; }
        if (ring.host) cudaFreeHost (ring.host);
        ring.host = nullptr; ring.slotCap = 0;
        const size_t cap = ((bytes * 2 < 65536 ? 65536 : bytes * 2) + 255) & ~(size_t) 255;
        ART_CUDA_CHECK (cudaHostAlloc ((void **) &ring.host, cap * 8, cudaHostAllocDefault));
        ring.slotCap = cap;
    }
    const unsigned int i = ring.next++ & 7u;
    if (ring.used[i])
        ART_CUDA_CHECK (cudaEventSynchronize (ring.done[i]));     // the copy issued eight launches ago
    else if (!ring.done[i])
        ART_CUDA_CHECK (cudaEventCreateWithFlags (&ring.done[i], cudaEventDisableTiming));
    char *slot = ring.host + (size_t) i * ring.slotCap;
    memcpy (slot, src, bytes);
    ART_CUDA_CHECK (cudaMemcpyAsync (d_dst, slot, bytes, cudaMemcpyHostToDevice, stream));
    ART_CUDA_CHECK (cudaEventRecord (ring.done[i], stream));
    ring.used[i] = true;
}

static void dispatch (ArtLaunchPlan &lp, std::vector<ArtJob> &jobs, int ctas, cudaStream_t stream, ArtDev *owner = nullptr)
{
    if (jobs.empty ())
        return;
    const int n = (int) jobs.size ();
    lp.k.numJobs = n;

    // tensor-core kernel on interleaved blocks of many channels: planar scratch (see above).  Segments of one call sit
    // next to each other in `jobs` and share their pointers: one scratch pair per call.
    artsample_t *scratch = nullptr;
    ArtXpose *d_xpose = nullptr;
    std::vector<ArtXpose> xin, xout;
    long long maxInFrames = 0, maxOutFrames = 0;
    const int C = lp.k.C;
    if (lp.umma && ctas > 0 && C >= 8 && lp.um.cg < 4) {            // (four channels per tile read 16-byte slices of the frames directly)
        bool ok = true;
        for (const ArtJob &j : jobs)
            ok &= j.inPlanes == nullptr && j.outPlanes == nullptr && j.inCS == 1 && j.inFS == C && j.outCS == 1 && j.outFS == C && j.prevAvail == 0;
        if (ok) {
            size_t floats = 0;
            std::vector<size_t> inAt, outAt;
            std::vector<int> groupOf (n, 0);                     // which call (scratch pair) a segment belongs to
            for (int i = 0; i < n; ) {
                int e = i;
                long long outFrames = 0, outFirst = -1, inFirst = (1LL << 62);    // inFirst may legitimately be negative (history)
                while (e < n && jobs[e].in == jobs[i].in && jobs[e].out == jobs[i].out) {
                    const long long end = (long long) jobs[e].nStart + jobs[e].outputs;
                    if (end > outFrames) outFrames = end;
                    if (jobs[e].outputs) {
                        // a piece of a pipelined host call starts at nStart > 0: frames before that belong to earlier pieces
                        // (already on their way to the host) and must neither be transposed in nor written back
                        if (outFirst < 0 || (long long) jobs[e].nStart < outFirst) outFirst = jobs[e].nStart;
                        const long long s0 = first_frame_read (jobs[e], lp.k, jobs[e].nStart);
                        if (s0 < inFirst) inFirst = s0;
                    }
                    if (jobs[e].histOut) {
                        const long long h0 = (long long) jobs[e].consumed - lp.k.T;
                        if (h0 < inFirst) inFirst = h0;
                    }
                    ++e;
                }
                if (outFirst < 0) outFirst = 0;
                inFirst = inFirst < 64 ? 0 : ((inFirst - 32) & ~31LL);            // a margin, and whole 128-byte lines
                const long long inFrames = jobs[i].inValid > 0 ? jobs[i].inValid : 0;
                if (inFirst > inFrames) inFirst = inFrames;
                const long long inPitch = (inFrames + 31) & ~31LL, outPitch = (outFrames + 31) & ~31LL;
                // src/dst are offset so that scratch index f still means frame f of the call
                xin.push_back ({ jobs[i].in + inFirst * C, reinterpret_cast<artsample_t *> (inFirst), inFrames - inFirst, inPitch });
                xout.push_back ({ reinterpret_cast<const artsample_t *> (outFirst), jobs[i].out + outFirst * C, outFrames - outFirst, outPitch });
                inAt.push_back (floats); floats += (size_t) C * inPitch;
                outAt.push_back (floats); floats += (size_t) C * outPitch;
                if (inFrames - inFirst > maxInFrames) maxInFrames = inFrames - inFirst;
                if (outFrames - outFirst > maxOutFrames) maxOutFrames = outFrames - outFirst;
                for (int q = i; q < e; ++q) groupOf[q] = (int) xin.size () - 1;
                i = e;
            }
            ART_CUDA_CHECK (cudaMallocAsync (&scratch, floats * sizeof (artsample_t), stream));
            for (int q = 0; q < n; ++q) {
                ArtJob &j = jobs[q];
                const int gi = groupOf[q];
                j.in = scratch + inAt[gi];   j.inFS = 1;  j.inCS = xin[gi].pitch;
                j.out = scratch + outAt[gi]; j.outFS = 1; j.outCS = xout[gi].pitch;
            }
            for (size_t gi = 0; gi < xin.size (); ++gi) {              // the offsets parked in dst / src above
                xin[gi].dst = scratch + inAt[gi] + reinterpret_cast<long long> (xin[gi].dst);
                xout[gi].src = scratch + outAt[gi] + reinterpret_cast<long long> (xout[gi].src);
            }
            const size_t ng = xin.size ();
            std::vector<ArtXpose> both (xin);
            both.insert (both.end (), xout.begin (), xout.end ());
            ART_CUDA_CHECK (cudaMallocAsync (&d_xpose, both.size () * sizeof (ArtXpose), stream));
            ART_CUDA_CHECK (cudaMemcpyAsync (d_xpose, both.data (), both.size () * sizeof (ArtXpose), cudaMemcpyHostToDevice, stream));
            if (maxInFrames > 0) {
                const dim3 grid ((unsigned int) ((maxInFrames + 31) / 32), (unsigned int) ((C + 31) / 32), (unsigned int) (ng < 65535 ? ng : 65535));
                art_deinterleave_kernel<<<grid, dim3 (32, 8), 0, stream>>> (d_xpose, C, (int) ng);
                ART_CUDA_CHECK (cudaGetLastError ());
                ++g_artLaunches;
            }
        }
    }
    bool anyHist = false;
    for (const ArtJob &j : jobs) anyHist |= j.histOut != nullptr;

    // periodic kernel: jobs whose filters coincide share one phase table.  The table depends on the
    // fractional read position of every phase, i.e. on (P, I, ratio, first output); contexts driven in
    // lock step (the usual many-stream case) collapse to a single table.
    int numTables = 0;
    if (lp.periodic && ctas > 0) {
        std::vector<int> reps;
        for (int i = 0; i < n; ++i) {
            int t = -1;
            for (size_t r = 0; r < reps.size () && t < 0; ++r) {
                const ArtJob &o = jobs[reps[r]];
                if (o.P == jobs[i].P && o.I == jobs[i].I && o.ratio == jobs[i].ratio &&
                    o.nStart == jobs[i].nStart && o.origin == jobs[i].origin)
                    t = (int) r;
                if (r >= 64) break;                     // keep the search linear; beyond that just add tables
            }
            if (t < 0) { t = (int) reps.size (); reps.push_back (i); }
            jobs[i].table = t;
        }
        numTables = (int) reps.size ();
        for (int t = 0; t < numTables; ++t)
            jobs[t].repJob = reps[t];
    }

    ArtJob *d_jobs = nullptr;
    if (n > 1) {
        ART_CUDA_CHECK (cudaMallocAsync (&d_jobs, sizeof (ArtJob) * n, stream));
        upload_jobs (d_jobs, jobs.data (), sizeof (ArtJob) * n, stream);
    }
    if (ctas > 0) {
        if (lp.periodic) {
            const size_t tableFloats = (size_t) numTables * lp.per.PB * lp.per.rowsPerCta * 8 * lp.per.Kp;
            const size_t tableInts = (size_t) n * lp.per.PB;
            const size_t bytes = lp.umma ? artUmmaTableBytes (lp.um, numTables, n, ctas)
                                          : tableFloats * sizeof (float) + tableInts * sizeof (int);
            void *tables = nullptr;
            const bool persistent = owner && n == 1;
            if (persistent) {
                if (bytes > owner->tableCap) {
                    ART_CUDA_CHECK (cudaStreamSynchronize (stream));
                    cudaFree (owner->tableBuf);
                    ART_CUDA_CHECK (cudaMalloc (&owner->tableBuf, bytes));
                    owner->tableCap = bytes;
                }
                tables = owner->tableBuf;
            }
            else
                ART_CUDA_CHECK (cudaMallocAsync (&tables, bytes, stream));
            if (lp.umma) {
                artUmmaCarve (lp.um, tables, numTables, n);
                // whole-frame copies in the converters need every job's block interleaved with the tile's channels adjacent and aligned
                bool vecIn = lp.um.cg > 1;
                for (const ArtJob &j : jobs)
                    vecIn = vecIn && j.inPlanes == nullptr && j.inCS == 1 && (j.inFS % lp.um.cg) == 0 &&
                            (reinterpret_cast<uintptr_t> (j.in) % (sizeof (float) * lp.um.cg)) == 0;
                artLaunchUmma (lp.k, lp.um, ctas, n, numTables, lp.smCount, jobs[0], d_jobs, vecIn, stream);
            }
            else {
                lp.per.Hblk = reinterpret_cast<float *> (tables);
                lp.per.S0 = reinterpret_cast<int *> (lp.per.Hblk + tableFloats);
                artLaunchPeriodic (lp.k, lp.per, lp.CV, ctas, n, numTables, jobs[0], d_jobs, stream);
            }
            ++g_pathLaunches[lp.umma ? 2 : 1];
            if (!persistent)
                ART_CUDA_CHECK (cudaFreeAsync (tables, stream));
            anyHist = false;                    // the periodic prep kernel moved the history already
        }
        else {
            lp.g.totalTiles = ctas;
            artLaunchGeneric (lp.k, lp.g, jobs[0], d_jobs, stream);
            ++g_pathLaunches[0];
        }
    }
    if (anyHist) {
        // usually one job carries the history update (the last block of an ASRC sequence, a single call): pass it by value
        int carriers = 0, which = 0;
        for (int i = 0; i < n; ++i) if (jobs[i].histOut) { ++carriers; which = i; }
        if (carriers == 1) artLaunchHistory (lp.k, jobs[which], nullptr, 1, stream);
        else artLaunchHistory (lp.k, jobs[0], d_jobs, n, stream);
    }
    if (scratch) {
        if (maxOutFrames > 0) {
            const size_t ng = xin.size ();
            const dim3 grid ((unsigned int) ((maxOutFrames + 31) / 32), (unsigned int) ((C + 31) / 32), (unsigned int) (ng < 65535 ? ng : 65535));
            art_interleave_kernel<<<grid, dim3 (32, 8), 0, stream>>> (d_xpose + ng, C, (int) ng);
            ART_CUDA_CHECK (cudaGetLastError ());
            ++g_artLaunches;
        }
        ART_CUDA_CHECK (cudaFreeAsync (d_xpose, stream));
        ART_CUDA_CHECK (cudaFreeAsync (scratch, stream));
    }
    if (d_jobs)
        ART_CUDA_CHECK (cudaFreeAsync (d_jobs, stream));
}

static void run_single (ArtDev *dev, const ArtCallPlan &p, ArtJob &job, cudaStream_t stream)
{
    ArtLaunchPlan lp;
    plan_launch (dev, p.st.ratio, p.st.ratio, true, p.outputs, p.outputs, true, lp);
    std::vector<ArtJob> jobs;
    const int ctas = append_job (lp, job, jobs, 0);
    dispatch (lp, jobs, ctas, stream, dev);
    finish_job (dev, p);
}

static void reserve (ArtDev *dev, size_t inFloats, size_t outFloats)
{
    if (inFloats > dev->inCap) {
        ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
        cudaFree (dev->d_in);
        dev->inCap = inFloats + inFloats / 4 + 1024;
        ART_CUDA_CHECK (cudaMalloc (&dev->d_in, dev->inCap * sizeof (artsample_t)));
    }
    if (outFloats > dev->outCap) {
        ART_CUDA_CHECK (cudaStreamSynchronize (dev->stream));
        cudaFree (dev->d_out);
        dev->outCap = outFloats + outFloats / 4 + 1024;
        ART_CUDA_CHECK (cudaMalloc (&dev->d_out, dev->outCap * sizeof (artsample_t)));
    }
}

/* ---- host-memory entry points ------------------------------------------------------------------------ */

/* One call, host memory, interleaved.  Large calls are cut into pieces of consecutive outputs; piece c
 * needs the input frames up to art_inputs_before(last output of c), so its upload, its kernels and its
 * download form a three-stage pipeline across pieces.  The samples are the same as for one big launch:
 * every output is evaluated from (P, I, n) alone. */
extern "C" int artDevRunHostInterleaved (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *in, artsample_t *out)
{
    ART_GUARD_BEGIN
    use_device (dev);
    const size_t C = dev->C;
    const size_t inFloats = (size_t) plan->inValid * C, outFloats = (size_t) plan->outputs * C;
    reserve (dev, inFloats, outFloats);

    // a piece costs ~25 us of launch latency; only cut when its copies take several times that
    static int pieceShift = -1;
    if (pieceShift < 0) { const char *e = getenv ("ART_B200_PIECE_SHIFT"); pieceShift = e ? atoi (e) : 20; }
    int pieces = (int) (((unsigned long long) plan->outputs * C) >> pieceShift);
    if (pieces < 1 || plan->pre) pieces = 1;
    if (pieces > 8) pieces = 8;
    if (getenv ("ART_B200_NOPIPE")) pieces = 1;

    if (pieces == 1) {
        if (inFloats)
            ART_CUDA_CHECK (cudaMemcpyAsync (dev->d_in, in, inFloats * sizeof (artsample_t), cudaMemcpyHostToDevice, dev->stream));
        ArtJob job;
        fill_job (dev, *plan, job);
        job.in = dev->d_in;  job.inFS = C;  job.inCS = 1;
        job.out = dev->d_out; job.outFS = C; job.outCS = 1;
        run_single (dev, *plan, job, dev->stream);
        if (outFloats)
            ART_CUDA_CHECK (cudaMemcpyAsync (out, dev->d_out, outFloats * sizeof (artsample_t), cudaMemcpyDeviceToHost, dev->stream));
        wait_for (dev, dev->stream, (inFloats + outFloats) * sizeof (artsample_t));
        return 0;
    }

    host_pipe_init (dev, 2 * (size_t) pieces);
    ArtLaunchPlan lp;
    plan_launch (dev, plan->st.ratio, plan->st.ratio, true, plan->outputs / pieces + 1, plan->outputs, true, lp);
    long long uploaded = 0;
    for (int c = 0; c < pieces; ++c) {
        const unsigned int n0 = (unsigned int) ((unsigned long long) plan->outputs * c / pieces);
        const unsigned int n1 = (unsigned int) ((unsigned long long) plan->outputs * (c + 1) / pieces);
        long long need = c == pieces - 1 ? (long long) plan->inValid : art_inputs_before (&plan->st, n1 - 1);
        if (need > plan->inValid) need = plan->inValid;
        if (need < uploaded) need = uploaded;
        if (need > uploaded)
            ART_CUDA_CHECK (cudaMemcpyAsync (dev->d_in + uploaded * C, in + uploaded * C, (size_t) (need - uploaded) * C * sizeof (artsample_t),
                                             cudaMemcpyHostToDevice, dev->sIn));
        uploaded = need;
        ART_CUDA_CHECK (cudaEventRecord (dev->events[2 * c], dev->sIn));
        ART_CUDA_CHECK (cudaStreamWaitEvent (dev->stream, dev->events[2 * c], 0));

        ArtJob job;
        fill_job (dev, *plan, job);
        job.in = dev->d_in;  job.inFS = C;  job.inCS = 1;
        job.out = dev->d_out; job.outFS = C; job.outCS = 1;
        job.nStart = n0;
        job.outputs = n1 - n0;
        job.inValid = (int) uploaded;
        if (c != pieces - 1) job.histOut = nullptr;          // the history moves once, after the last upload
        std::vector<ArtJob> jobs;
        const int ctas = append_job (lp, job, jobs, 0);
        dispatch (lp, jobs, ctas, dev->stream, dev);

        ART_CUDA_CHECK (cudaEventRecord (dev->events[2 * c + 1], dev->stream));
        ART_CUDA_CHECK (cudaStreamWaitEvent (dev->sOut, dev->events[2 * c + 1], 0));
        if (n1 > n0)
            ART_CUDA_CHECK (cudaMemcpyAsync (out + (size_t) n0 * C, dev->d_out + (size_t) n0 * C, (size_t) (n1 - n0) * C * sizeof (artsample_t),
                                             cudaMemcpyDeviceToHost, dev->sOut));
    }
    finish_job (dev, *plan);
    // the last download was enqueued after everything else of this call (sOut waits for the last kernel)
    wait_for (dev, dev->sOut, (inFloats + outFloats) * sizeof (artsample_t));
    return 0;
    ART_GUARD_END (-1)
}

/* Many contexts, host memory, interleaved: the same three-stage pipeline over groups of contexts -- a
 * group's uploads, one launch for the group, its downloads (group size 1 measured best on B200: 64 us
 * per 2 MB stereo stream against a PCIe floor of 42 us). */
extern "C" int artDevRunBatchInterleaved (ArtDev *const *devs, const ArtCallPlan *plans, int count,
                                           const artsample_t *const *d_in, artsample_t *const *d_out, void *stream);

extern "C" int artDevRunHostBatchInterleaved (ArtDev *const *devs, const ArtCallPlan *plans, int count,
                                               const artsample_t *const *in, artsample_t *const *out)
{
    ART_GUARD_BEGIN
    if (count <= 0) return 0;
    ArtDev *lead = devs[0];
    use_device (lead);
    const int group = 1;                   // measured: larger groups coarsen the pipeline more than they save
    const int groups = (count + group - 1) / group;
    host_pipe_init (lead, 2 * (size_t) groups);
    const size_t C = lead->C;
    std::vector<const artsample_t *> dIn (count);
    std::vector<artsample_t *> dOut (count);
    for (int g = 0; g < groups; ++g) {
        const int i0 = g * group, i1 = i0 + group < count ? i0 + group : count;
        for (int i = i0; i < i1; ++i) {
            ArtDev *dev = devs[i];
            const ArtCallPlan &p = plans[i];
            if (dev->device != lead->device)
                artRaise ("a batch must live on one GPU");
            const size_t inFloats = (size_t) p.inValid * C, outFloats = (size_t) p.outputs * C;
            reserve (dev, inFloats, outFloats);
            dIn[i] = dev->d_in;
            dOut[i] = dev->d_out;
            if (inFloats)
                ART_CUDA_CHECK (cudaMemcpyAsync (dev->d_in, in[i], inFloats * sizeof (artsample_t), cudaMemcpyHostToDevice, lead->sIn));
        }
        ART_CUDA_CHECK (cudaEventRecord (lead->events[2 * g], lead->sIn));
        ART_CUDA_CHECK (cudaStreamWaitEvent (lead->stream, lead->events[2 * g], 0));
        if (artDevRunBatchInterleaved (devs + i0, plans + i0, i1 - i0, dIn.data () + i0, dOut.data () + i0, lead->stream))
            throw ArtError { -1 };
        ART_CUDA_CHECK (cudaEventRecord (lead->events[2 * g + 1], lead->stream));
        ART_CUDA_CHECK (cudaStreamWaitEvent (lead->sOut, lead->events[2 * g + 1], 0));
        for (int i = i0; i < i1; ++i) {
            const size_t outFloats = (size_t) plans[i].outputs * C;
            if (outFloats)
                ART_CUDA_CHECK (cudaMemcpyAsync (out[i], devs[i]->d_out, outFloats * sizeof (artsample_t), cudaMemcpyDeviceToHost, lead->sOut));
        }
    }
    // sOut's last download waits for the last launch, which waits for the last upload
    wait_for (lead, lead->sOut, (size_t) 1 << 20);
    ART_CUDA_CHECK (cudaStreamSynchronize (lead->stream));      // (already idle: only the history kernels of the last group can trail)
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevRunHostPlanar (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *const *in, artsample_t *const *out)
{
    ART_GUARD_BEGIN
    use_device (dev);
    const size_t C = dev->C;
    const size_t nin = plan->inValid, nout = plan->outputs;
    reserve (dev, nin * C, nout * C);
    for (size_t c = 0; c < C && nin; ++c)
        ART_CUDA_CHECK (cudaMemcpyAsync (dev->d_in + c * nin, in[c], nin * sizeof (artsample_t), cudaMemcpyHostToDevice, dev->stream));
    ArtJob job;
    fill_job (dev, *plan, job);
    job.in = dev->d_in;  job.inFS = 1;  job.inCS = (long long) nin;
    job.out = dev->d_out; job.outFS = 1; job.outCS = (long long) nout;
    run_single (dev, *plan, job, dev->stream);
    for (size_t c = 0; c < C && nout; ++c)
        ART_CUDA_CHECK (cudaMemcpyAsync (out[c], dev->d_out + c * nout, nout * sizeof (artsample_t), cudaMemcpyDeviceToHost, dev->stream));
    wait_for (dev, dev->stream, (nin + nout) * C * sizeof (artsample_t));
    return 0;
    ART_GUARD_END (-1)
}

/* ---- device-memory entry points ------------------------------------------------------------------------ */

extern "C" int artDevRunDeviceInterleaved (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *d_in, artsample_t *d_out, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    dev->lastStream = st;
    ArtJob job;
    fill_job (dev, *plan, job);
    job.in = d_in;   job.inFS = dev->C;  job.inCS = 1;
    job.out = d_out; job.outFS = dev->C; job.outCS = 1;
    run_single (dev, *plan, job, st);
    return 0;
    ART_GUARD_END (-1)
}

extern "C" int artDevRunDevicePlanar (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *const *d_in, artsample_t *const *d_out, void *stream)
{
    ART_GUARD_BEGIN
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    dev->lastStream = st;
    const int C = dev->C;
    ArtJob job;
    fill_job (dev, *plan, job);
    job.inFS = 1; job.outFS = 1;

    // equally spaced planes need no pointer table
    bool uniformIn = d_in != nullptr, uniformOut = true;
    const long long pin = (C > 1 && d_in) ? (long long) (d_in[1] - d_in[0]) : 0;
    const long long pout = C > 1 ? (long long) (d_out[1] - d_out[0]) : 0;
    for (int c = 1; c < C; ++c) {
        if (d_in && d_in[c] - d_in[c - 1] != pin) uniformIn = false;
        if (d_out[c] - d_out[c - 1] != pout) uniformOut = false;
    }
    void *table = nullptr;
    if (uniformIn || !d_in) { job.in = d_in ? d_in[0] : nullptr; job.inCS = pin; }
    if (uniformOut) { job.out = d_out[0]; job.outCS = pout; }
    if ((d_in && !uniformIn) || !uniformOut) {
        std::vector<const void *> h (2 * (size_t) C, nullptr);
        for (int c = 0; c < C; ++c) { h[c] = d_in ? d_in[c] : nullptr; h[C + c] = d_out[c]; }
        ART_CUDA_CHECK (cudaMallocAsync (&table, sizeof (void *) * 2 * C, st));
        // pageable source: staged by the runtime before the call returns
        ART_CUDA_CHECK (cudaMemcpyAsync (table, h.data (), sizeof (void *) * 2 * C, cudaMemcpyHostToDevice, st));
        if (d_in && !uniformIn) job.inPlanes = reinterpret_cast<const artsample_t *const *> (table);
        if (!uniformOut) job.outPlanes = reinterpret_cast<artsample_t *const *> (table) + C;
    }
    run_single (dev, *plan, job, st);
    if (table)
        ART_CUDA_CHECK (cudaFreeAsync (table, st));
    return 0;
    ART_GUARD_END (-1)
}

/* ---- many contexts, one launch ----------------------------------------------------------------------------- */

extern "C" int artDevRunBatchInterleaved (ArtDev *const *devs, const ArtCallPlan *plans, int count,
                                           const artsample_t *const *d_in, artsample_t *const *d_out, void *stream)
{
    ART_GUARD_BEGIN
    if (count <= 0) return 0;
    ArtDev *lead = devs[0];
    use_device (lead);
    cudaStream_t st = stream ? (cudaStream_t) stream : lead->stream;

    double minRatio = plans[0].st.ratio, maxRatio = plans[0].st.ratio;
    bool oneRatio = true;
    unsigned int maxOut = 0;
    unsigned long long totalOut = 0;
    for (int i = 0; i < count; ++i) {
        if (devs[i]->bank != lead->bank || devs[i]->C != lead->C || devs[i]->mode != lead->mode || devs[i]->device != lead->device)
            artRaise ("a batch must hold contexts of one configuration on one GPU");
        if (plans[i].st.ratio != plans[0].st.ratio) oneRatio = false;
        if (plans[i].st.ratio < minRatio) minRatio = plans[i].st.ratio;
        if (plans[i].st.ratio > maxRatio) maxRatio = plans[i].st.ratio;
        if (plans[i].outputs > maxOut) maxOut = plans[i].outputs;
        totalOut += plans[i].outputs;
    }

    ArtLaunchPlan lp;
    plan_launch (lead, minRatio, maxRatio, oneRatio, maxOut, totalOut, true, lp);
    std::vector<ArtJob> jobs;
    jobs.reserve (count);
    int ctas = 0;
    for (int i = 0; i < count; ++i) {
        ArtJob j;
        fill_job (devs[i], plans[i], j);
        j.in = d_in ? d_in[i] : nullptr;  j.inFS = lead->C;  j.inCS = 1;
        j.out = d_out[i];                 j.outFS = lead->C; j.outCS = 1;
        devs[i]->lastStream = st;
        ctas += append_job (lp, j, jobs, ctas);
    }
    dispatch (lp, jobs, ctas, st, count == 1 ? lead : nullptr);
    for (int i = 0; i < count; ++i)
        finish_job (devs[i], plans[i]);
    return 0;
    ART_GUARD_END (-1)
}

/* ---- consecutive blocks of one stream, one launch (ASRC) ------------------------------------------------------ */

extern "C" int artDevRunBlocksInterleaved (ArtDev *dev, const ArtCallPlan *plans, int count,
                                            const long long *inOffset, const long long *outOffset,
                                            const artsample_t *d_in, artsample_t *d_out, void *stream)
{
    ART_GUARD_BEGIN
    if (count <= 0) return 0;
    use_device (dev);
    cudaStream_t st = stream ? (cudaStream_t) stream : dev->stream;
    dev->lastStream = st;
    const int C = dev->C;

    double minRatio = plans[0].st.ratio, maxRatio = plans[0].st.ratio;
    unsigned int maxOut = 0;
    unsigned long long totalOut = 0;
    long long totalIn = 0;
    for (int i = 0; i < count; ++i) {
        if (plans[i].st.ratio < minRatio) minRatio = plans[i].st.ratio;
        if (plans[i].st.ratio > maxRatio) maxRatio = plans[i].st.ratio;
        if (plans[i].outputs > maxOut) maxOut = plans[i].outputs;
        totalOut += plans[i].outputs;
        totalIn += plans[i].consumed;
    }
    ArtLaunchPlan lp;
    plan_launch (dev, minRatio, maxRatio, false, maxOut, totalOut, false, lp);

    std::vector<ArtJob> jobs;
    jobs.reserve (count + 1);
    int ctas = 0;
    for (int i = 0; i < count; ++i) {
        ArtJob j;
        fill_job (dev, plans[i], j);
        j.histOut = nullptr;
        j.prevAvail = inOffset[i];
        j.in = d_in + inOffset[i] * C;    j.inFS = C;  j.inCS = 1;
        j.out = d_out + outOffset[i] * C; j.outFS = C; j.outCS = 1;
        ctas += append_job (lp, j, jobs, ctas);
    }
    // the history after the sequence: newest T frames of (history ++ all consumed input)
    if (totalIn) {
        ArtJob hj;
        memset (&hj, 0, sizeof hj);
        hj.hist = dev->hist[dev->cur];
        hj.histOut = dev->hist[dev->cur ^ 1];
        hj.in = d_in; hj.inFS = C; hj.inCS = 1;
        hj.inValid = (int) totalIn;
        hj.consumed = totalIn;
        hj.tile0 = ctas;                    // owns no tiles: outputs == 0
        jobs.push_back (hj);
    }
    dispatch (lp, jobs, ctas, st);
    if (totalIn)
        dev->cur ^= 1;
    return 0;
    ART_GUARD_END (-1)
}
