// Multi-GPU dispatch behind the C ABI: one vocoder handle, one host thread and one CUDA context use per GPU.
//
// The reference vocodes one utterance at a time in one process (/root/reference src/lib.rs:83-104, 110-159; its
// author notes the sentences could run in parallel, src/phonemes.rs:677-680).  Utterances are independent, so a
// batch shards across the GPUs of a box with NO data-path exchange: xdtts_pool_infer_batch
// assigns utterances longest-first to the least-loaded device (the rule of xdtts_b200/shard.py, ties to the lowest
// device), hands every device its sub-batch on its own worker thread and returns when all are done.  An utterance
// draws the phase stream of its position in the CALLER's batch and every device of one call uses the same seed, so the
// result does not depend on the number of devices (bitwise, when opts.run_frames fixes the run split).
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "api_internal.h"

using namespace xdtts;
#define fail xdtts::set_error

struct xdtts_pool {
    struct Worker {
        xdtts_gl* gl = nullptr;
        int device = 0;
        std::thread th;
        // job of the current call
        std::vector<const float*> ins, phases;
        std::vector<float*> outs;
        std::vector<int> Ts, ids;
        int rc = 0;
        std::string err;
        bool has_job = false, done = false;
    };
    std::vector<Worker*> workers;
    std::mutex mu;                 // protects the job hand-over
    std::condition_variable cv_job, cv_done;
    std::mutex call_mu;            // one pooled call at a time
    bool stop = false;
    int kind = 0;                  // 0 mels, 1 magnitudes (of the current call)
    unsigned long long seed = 0, calls = 0;
    bool fixed_seed = false;
    unsigned long long call_seed = 0;
    bool use_phase = false;
};

// deterministic on every caller: longest first, to the least-loaded device, ties to the lowest device index
static void lpt_assign(const int* Ts, int B, int n_dev, int* dev_of) {
    std::vector<int> order(B);
    for (int i = 0; i < B; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return Ts[a] > Ts[b]; });
    std::vector<long long> load(n_dev, 0);
    for (int i : order) {
        int best = 0;
        for (int d = 1; d < n_dev; d++)
            if (load[d] < load[best]) best = d;
        load[best] += Ts[i];
        dev_of[i] = best;
    }
}

static void worker_main(xdtts_pool* p, xdtts_pool::Worker* w) {
    cudaSetDevice(w->device);
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->stop || w->has_job; });
            if (p->stop) return;
        }
        int rc = XDTTS_OK;
        if (!w->Ts.empty())
            rc = gl_batch_common(w->gl, p->kind, w->ins.data(), w->Ts.data(), (int)w->Ts.size(), p->use_phase ? w->phases.data() : nullptr,
                                 w->outs.data(), nullptr, &p->call_seed, w->ids.data());
        {
            std::lock_guard<std::mutex> lk(p->mu);
            w->rc = rc;
            if (rc) w->err = xdtts_last_error();   // the message is thread-local: carry it to the caller's thread
            w->has_job = false;
            w->done = true;
        }
        p->cv_done.notify_all();
    }
}

// host-only: the assignment rule by itself (slot index 0..n_slots-1 per utterance)
extern "C" int xdtts_shard_assign(const int* Ts, int B, int n_slots, int* slot_of_utt) {
    if (!Ts || !slot_of_utt || B < 1 || n_slots < 1) return fail(XDTTS_ERR_BAD_ARG, "shard_assign: bad argument");
    lpt_assign(Ts, B, n_slots, slot_of_utt);
    return XDTTS_OK;
}

extern "C" void xdtts_pool_destroy(xdtts_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto* w : p->workers) {
        if (w->th.joinable()) w->th.join();
        if (w->gl) xdtts_gl_destroy(w->gl);
        delete w;
    }
    delete p;
}

extern "C" int xdtts_pool_create(const float* mel_basis, int n_mels, int K, int noverlap, float power, int n_iter, float momentum,
                                 const xdtts_gl_opts* opts, const int* devices, int n_devices, xdtts_pool** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "pool_create: out is null");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(XDTTS_ERR_CUDA, "pool_create: no CUDA device (this library has no CPU path)");
    }
    if (n_devices < 0 || n_devices > 64 || (n_devices > 0 && !devices)) return fail(XDTTS_ERR_BAD_ARG, "pool_create: bad device list");
    std::vector<int> devs;
    if (n_devices == 0)
        for (int d = 0; d < n_dev; d++) devs.push_back(d);   // every visible device
    else
        devs.assign(devices, devices + n_devices);
    for (size_t i = 0; i < devs.size(); i++) {
        if (devs[i] < 0 || devs[i] >= n_dev) return fail(XDTTS_ERR_BAD_ARG, "pool_create: device %d of %d", devs[i], n_dev);
        for (size_t j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail(XDTTS_ERR_BAD_ARG, "pool_create: device %d listed twice", devs[i]);
    }
    xdtts_pool* p = new (std::nothrow) xdtts_pool();
    if (!p) return fail(XDTTS_ERR_OOM, "pool_create: out of host memory");
    p->seed = opts ? opts->seed : 0;
    p->fixed_seed = opts && opts->fixed_seed;
    for (int d : devs) {
        xdtts_pool::Worker* w = new (std::nothrow) xdtts_pool::Worker();
        if (!w) {
            xdtts_pool_destroy(p);
            return fail(XDTTS_ERR_OOM, "pool_create: out of host memory");
        }
        w->device = d;
        p->workers.push_back(w);
        int rc = xdtts_gl_create(mel_basis, n_mels, K, noverlap, power, n_iter, momentum, opts, d, &w->gl);
        if (rc) {
            xdtts_pool_destroy(p);
            return rc;   // message already set by gl_create
        }
    }
    for (auto* w : p->workers) w->th = std::thread(worker_main, p, w);
    *out = p;
    return XDTTS_OK;
}

extern "C" int xdtts_pool_n_devices(const xdtts_pool* p) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "pool_n_devices: pool is null");
    return (int)p->workers.size();
}

extern "C" int xdtts_pool_out_len(const xdtts_pool* p, int T) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "pool_out_len: pool is null");
    return xdtts_gl_out_len(p->workers[0]->gl, T);
}

extern "C" int xdtts_pool_assignment(const xdtts_pool* p, const int* Ts, int B, int* device_of_utt) {
    if (!p || !Ts || !device_of_utt || B < 1) return fail(XDTTS_ERR_BAD_ARG, "pool_assignment: bad argument");
    std::vector<int> slot(B);
    lpt_assign(Ts, B, (int)p->workers.size(), slot.data());
    for (int i = 0; i < B; i++) device_of_utt[i] = p->workers[slot[i]]->device;
    return XDTTS_OK;
}

static int pool_run(xdtts_pool* p, int kind, const float* const* ins, const int* Ts, int B, const float* const* phases,
                    float* const* outs) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "pool_infer: pool is null");
    if (!ins || !Ts || !outs) return fail(XDTTS_ERR_BAD_ARG, "pool_infer: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "pool_infer: B = %d", B);
    for (int b = 0; b < B; b++) {
        if (!ins[b] || !outs[b] || (phases && !phases[b])) return fail(XDTTS_ERR_BAD_ARG, "pool_infer: null buffer for utterance %d", b);
        if (Ts[b] < 2) return fail(XDTTS_ERR_SHAPE, "pool_infer: utterance %d has T = %d frames, need >= 2", b, Ts[b]);
    }
    std::lock_guard<std::mutex> call(p->call_mu);
    const int nd = (int)p->workers.size();
    std::vector<int> slot(B);
    lpt_assign(Ts, B, nd, slot.data());
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->kind = kind;
        p->use_phase = phases != nullptr;
        if (!phases) p->call_seed = p->seed + (p->fixed_seed ? 0ull : p->calls++) * 0xD1B54A32D192ED03ull;
        for (auto* w : p->workers) {
            w->ins.clear(); w->phases.clear(); w->outs.clear(); w->Ts.clear(); w->ids.clear();
            w->rc = 0; w->err.clear(); w->done = false;
        }
        for (int b = 0; b < B; b++) {   // ascending utterance index within a device
            xdtts_pool::Worker* w = p->workers[slot[b]];
            w->ins.push_back(ins[b]);
            if (phases) w->phases.push_back(phases[b]);
            w->outs.push_back(outs[b]);
            w->Ts.push_back(Ts[b]);
            w->ids.push_back(b);
        }
        for (auto* w : p->workers) w->has_job = true;
    }
    p->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->cv_done.wait(lk, [&] {
            for (auto* w : p->workers)
                if (!w->done) return false;
            return true;
        });
    }
    for (auto* w : p->workers)
        if (w->rc) return fail(w->rc, "pool_infer: device %d: %s", w->device, w->err.c_str());
    return XDTTS_OK;
}

extern "C" int xdtts_pool_infer_batch(xdtts_pool* p, const float* const* mels, const int* Ts, int B,
                                      const float* const* init_phases_or_null, float* const* outs) {
    return pool_run(p, 0, mels, Ts, B, init_phases_or_null, outs);
}

extern "C" int xdtts_pool_from_mag_batch(xdtts_pool* p, const float* const* mags, const int* Ts, int B,
                                         const float* const* init_phases_or_null, float* const* outs) {
    return pool_run(p, 1, mags, Ts, B, init_phases_or_null, outs);
}
