// Engine: static launch plan for one UNet velocity evaluation + the Euler integrator + the C ABI (include/rfv.h).
//
// The plan is built once per (architecture, resolution, micro-batch): every activation buffer, GroupNorm statistics
// slice, packed weight and TMA tensor map is allocated / encoded at rfv_create, so the hot path makes no allocation
// and no host decision beyond iterating a vector of launches.  Wiring follows UNet.forward (models/unet.py:229-275).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <tuple>
#include <memory>
#include <string>
#include <vector>

#include "attn.cuh"
#include "attn_umma.cuh"
#include "common.cuh"
#include "conv_mma.cuh"
#include "conv_params.h"
#include "conv_wa.cuh"
#include "conv_umma.cuh"
#include "conv_umma2.cuh"
#include "metrics.cuh"
#include "misc_kernels.cuh"
#include "rfv.h"
#include "train_kernels.cuh"
#include "wgrad.cuh"

#define RFV_EXPORT extern "C" __attribute__((visibility("default")))

namespace rfv {

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU_CHECK(expr)                                                                                              \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess)                                                                                      \
            return fail(RFV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
    } while (0)
#define RFV_TRY(expr)          \
    do {                       \
        int _rc = (expr);      \
        if (_rc != 0) return _rc; \
    } while (0)

static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

// One launch of a forward-path kernel; pdl = with the programmatic-stream-serialization attribute (common.cuh: pdl_wait).
template <typename... KA, typename... A>
static cudaError_t klaunch(bool pdl, void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, k, std::forward<A>(args)...);
}
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// ---------------------------------------------------------------------------------------------------------
// device memory bookkeeping
// ---------------------------------------------------------------------------------------------------------
struct Act {  // NHWC bf16 activation of the current micro-batch
    bf16* p = nullptr;
    int C = 0, H = 0, W = 0;
    float* stats = nullptr;  // [cap][C/slab][2]
    size_t bytes = 0;
    int refs = 0;
    bf16* grad = nullptr;    // training: dL/d(this tensor), same layout
    int consumers = 0;       // training: number of differentiable consumers (the last one, in forward order, runs first
                             // in the backward pass and overwrites `grad`; the others accumulate)
};
typedef std::shared_ptr<Act> ActP;

struct Param {  // one reference state_dict tensor
    std::string name;
    int64_t numel = 0;
    float* f32 = nullptr;  // device copy in reference layout
    bool loaded = false;
    std::function<int(cudaStream_t)> repack;  // refresh derived buffers after upload
    std::function<int(float*, cudaStream_t)> readback;  // optional: reconstruct fp32 from the packed form
    int64_t goff = 0;        // training: offset of this tensor's slot in the flat gradient / Adam-moment buffers
    int final_block = 1 << 30;   // training: backward block after which this tensor's gradient is final (blocks run back to front)
    int O = 0, I = 0, KK = 0;  // conv weights (KK > 1): the slot is laid out [O][KK][I] (what the wgrad kernel writes)
    float* bound = nullptr;  // caller-owned fp32 storage (torch Parameter) the optimizer also writes
};

struct RunCtx {
    int B = 0;               // images in this micro-batch
    const float* x = nullptr;    // [B,C,S,S] fp32 NCHW input state
    const float* x1 = nullptr;   // if set: input is (1-t) x + t x1
    const float* t = nullptr;    // [B] or nullptr (uniform t_scalar)
    float t_scalar = 0.f;
    float* out = nullptr;    // v (mode 0) or x in place (mode 1); unused in mode 2
    float* traj = nullptr;
    const float* tgt_x0 = nullptr;  // target = tgt_x1 - tgt_x0 for the MSE accumulator
    const float* tgt_x1 = nullptr;
    float* mse = nullptr;
    int mode = 0;
    float dt = 0.f;
    // training only
    bool train = false;
    uint32_t drop_thresh = 0, seed = 0;
    float drop_scale = 1.f;
    // Euler loops: the time projections of ALL steps are computed once up front (t_i = i * dt is known before the loop), into
    // rows 0 .. num_steps-1 of the projection table; step i then reads row temb_row and skips the time-MLP kernels.  -1: off.
    int temb_row = -1;
    bool temb_only = false;   // run just the time-MLP ops (the table fill)
    bool pdl = false;         // launch with programmatic dependent launch (set by run_forward: sampling engines, not while profiling)
};

struct Op {
    std::string label;   // e.g. "conv:enc_blocks.0.conv1"
    std::string kind;    // kernel class for the profile report
    double flops = 0;    // algorithmic FLOPs per image
    double bytes = 0;    // algorithmic HBM bytes per image (HBM-bound kernels: GroupNorm apply / backward), else 0
    std::function<cudaError_t(const RunCtx&, cudaStream_t)> run;
    int lane = 0;        // backward pass: 0 = main stream, 1 = side stream (weight / bias gradients)
    int sync = 0;        // 1: the side stream first waits for everything enqueued on main; 2: main first waits for side
};

struct ConvLayer {
    std::string name;
    int C0 = 0, Cout = 0, ks = 3, stride = 1, ups = 0, C1a = 0, C1b = 0, K0 = 0, Ktot = 0;
    bool subpixel = false;  // ups layer packed as 4 phases x (2x2 taps): w is [4*Cout][4*C0]
    bf16* w = nullptr;
    float* bias = nullptr;  // fused (conv bias + shortcut bias)
    int iw = -1, ib = -1, isw = -1, isb = -1;  // parameter indices (weight, bias, shortcut weight, shortcut bias)
    bf16* wa = nullptr;                 // weights-as-A block layout (conv_wa.cuh), packed from `w` after every repack
    std::vector<int> wparams;           // parameters whose repack rewrites `w`
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace rfv

using namespace rfv;

struct rfv_engine {
    rfv_config cfg{};
    int num_sms = 148;
    int cap = 0;       // micro-batch capacity (even)
    int slab_shift = 3;
    int td = 256, sumC = 0;
    bool keep_acts = false, use_umma = true, use_wa = true;
    int fuse_mode = 1;   // GroupNorm+SiLU inside the consuming conv: 0 never, 1 where measured faster, 2 wherever the kernel applies
    int cluster = 1;  // CTAs per cluster for weight multicast (flags bits 8-10 select 2 or 4; measured slower than 1 on B200)
    EncodeTiledFn encode = nullptr;

    std::vector<void*> allocs;
    std::vector<Param> params;
    std::map<std::string, int> param_index;
    std::vector<std::unique_ptr<ConvLayer>> convs;
    std::vector<Op> ops;
    std::map<std::string, ActP> named_acts;
    std::multimap<size_t, bf16*> free_pool;

    float* stats_arena = nullptr;
    size_t stats_floats = 0, stats_used = 0;
    float* temb_act = nullptr;   // [cap][td]
    float* tproj = nullptr;      // [cap][sumC]
    float* t_steps = nullptr;    // [cap] step times of the current Euler loop
    float* wcat = nullptr;       // [sumC][td]
    float* bcat = nullptr;       // [sumC]
    float* scratch_x = nullptr;  // [cap][C][S][S] fp32 (straightness state)
    float* xbuf[2] = {nullptr, nullptr};  // host-path double buffers
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_cmp = nullptr;
    cudaEvent_t ev_in[2]{}, ev_done[2]{}, ev_out[2]{};
    cudaEvent_t ev_weights = nullptr;  // last parameter upload (made on the caller's stream)
    cudaEvent_t ev_last = nullptr;     // end of the last enqueued call: the arena is shared, calls on different
    cudaStream_t last_stream = nullptr;  // streams are chained through this event
    bool have_last = false;

    int enter(cudaStream_t s) {
        if (have_last && s != last_stream) {
            cudaError_t e = cudaStreamWaitEvent(s, ev_last, 0);
            if (e != cudaSuccess) return fail(RFV_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    int leave(cudaStream_t s) {
        cudaError_t e = cudaEventRecord(ev_last, s);
        if (e != cudaSuccess) return fail(RFV_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
        last_stream = s;
        have_last = true;
        return 0;
    }

    // ----- training state (RFV_FLAG_TRAIN) ---------------------------------------------------------------------
    bool train = false;
    std::vector<Op>* rec = &ops;                 // list that push() appends to
    std::vector<std::vector<Op>> bwd_blocks;     // one per forward stage, executed back to front
    float *gflat = nullptr, *mflat = nullptr, *vflat = nullptr;  // flat gradient / Adam moments, slot order = parameter order
    int64_t gtotal = 0;
    float* cs_arena = nullptr;                   // GroupNorm-backward per-(image, channel) sums, one slice per norm site
    size_t cs_floats = 0, cs_used = 0;
    bf16* scratch[3] = {nullptr, nullptr, nullptr};
    size_t scratch_elems = 0;                    // per image
    float *d_tproj = nullptr, *temb_emb = nullptr, *temb_z1 = nullptr, *temb_h1 = nullptr, *temb_z2 = nullptr;
    float *d_tz2 = nullptr, *d_tz1 = nullptr;
    float *dv_buf = nullptr, *lse = nullptr, *delta = nullptr, *zero_bias = nullptr, *norm2 = nullptr, *norm_partial = nullptr;
    AdamSeg* d_segs = nullptr;
    int2* d_adam_blocks = nullptr;
    int n_adam_blocks = 0;
    bool adam_dirty = true;
    cudaGraphExec_t repack_graph = nullptr;
    cudaStream_t s_bwd = nullptr, s_side = nullptr;
    cudaEvent_t ev_bwd[4]{};
    // Second sampling lane: batches beyond one micro-batch are integrated as TWO independent chains (this engine and a twin with
    // its own arena and weight copies) on two streams, enqueued step by step from the calling thread, so one chain's HBM-bound
    // kernels (GroupNorm apply, thin convs) run under the other chain's tensor-bound convolutions.  Created on first use.
    rfv_engine* lane = nullptr;
    bool use_lanes = true;     // RFV_FLAG_ONE_LANE clears it
    // Whole-loop executor: the N-step Euler loop of one micro-batch (time-table fill + N x ~69 launches on fixed engine-owned
    // buffers) is captured once per (rows, steps, state buffer) and replayed as ONE graph launch.
    struct LoopGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
    std::map<std::tuple<int, int, int>, LoopGraph> loop_graphs;
    bool use_graphs = true;    // RFV_FLAG_NO_GRAPH clears it
    bool use_pdl = true;       // RFV_FLAG_NO_PDL clears it: kernels of a sampling chain launched as programmatic dependents
    cudaEvent_t ev_fork = nullptr, ev_join[2]{};
    // Gradient buckets for a data-parallel caller: slots are laid out in the order their gradients become final during the
    // backward pass, and cut into a few contiguous ranges; run_backward records, per bucket, an event on each backward stream
    // once the last block that writes into the range has been enqueued, so the caller can all-reduce a finished range while
    // the rest of the backward pass still runs.
    struct GradBucket { int64_t off = 0, numel = 0; int ready_block = 0; cudaEvent_t ev[2] = {nullptr, nullptr}; };
    std::vector<GradBucket> buckets;
    int temb_block = 0;
    bool buckets_valid = false;   // the events belong to the last backward pass of the last rfv_train_accumulate
    RunCtx fwd_rc;             // rfv_train_forward's context, replayed by rfv_train_backward
    bool have_fwd = false;
    bool two_streams = true;   // RFV_FLAG_ONE_STREAM: run the whole backward pass on one stream (A/B testing)
    int norm_sites = 0;
    struct TimeProj { int off, Cout, iw, ib, icb; };
    std::vector<TimeProj> time_projs;
    struct BwdProf { const Op* op; cudaEvent_t e0, e1; };
    std::vector<BwdProf> bwd_prof;

    int64_t launches = 0;
    double flops_per_image = 0;
    bool profiling = false;
    std::map<std::string, std::pair<double, int64_t>> prof;  // kind -> (ms, launches)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<std::string> prof_kinds;

    ~rfv_engine() {
        delete lane;
        for (auto& b : buckets) for (auto& e : b.ev) if (e) cudaEventDestroy(e);
        for (auto& kv : loop_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& e : ev_join) if (e) cudaEventDestroy(e);
        for (void* p : allocs) cudaFree(p);
        for (int i = 0; i < 2; ++i) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
        if (repack_graph) cudaGraphExecDestroy(repack_graph);
        for (auto& e : ev_bwd) if (e) cudaEventDestroy(e);
        if (s_bwd) cudaStreamDestroy(s_bwd);
        if (s_side) cudaStreamDestroy(s_side);
        if (ev_weights) cudaEventDestroy(ev_weights);
        if (ev_last) cudaEventDestroy(ev_last);
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        if (s_cmp) cudaStreamDestroy(s_cmp);
        for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    }

    // ----- allocation -----------------------------------------------------------------------------------
    template <typename T>
    int dalloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256));
        if (e != cudaSuccess) return fail(RFV_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        allocs.push_back(p);
        *out = reinterpret_cast<T*>(p);
        return 0;
    }
    int new_act(ActP* out, int C, int H, int W, bool with_stats) {
        auto a = std::make_shared<Act>();
        a->C = C; a->H = H; a->W = W; a->refs = 1;
        a->bytes = (size_t)cap * H * W * C * sizeof(bf16);
        auto it = keep_acts ? free_pool.end() : free_pool.find(a->bytes);
        if (it != free_pool.end()) { a->p = it->second; free_pool.erase(it); }
        else RFV_TRY(dalloc(&a->p, a->bytes / sizeof(bf16)));
        if (with_stats) {
            size_t n = (size_t)cap * (C >> slab_shift) * 2;
            if (stats_used + n > stats_floats) return fail(RFV_ERR_NOMEM, "stats arena exhausted");
            a->stats = stats_arena + stats_used;
            stats_used += n;
        }
        *out = a;
        return 0;
    }
    void release(ActP& a) {
        if (!a) return;
        if (--a->refs == 0 && !keep_acts) free_pool.insert({a->bytes, a->p});
    }
    static void retain(ActP& a) { a->refs++; }

    // ----- parameters -----------------------------------------------------------------------------------
    int add_param(const std::string& name, int64_t numel, int* idx) {
        Param p;
        p.name = "velocity_net." + name;
        p.numel = numel;
        RFV_TRY(dalloc(&p.f32, (size_t)numel));
        p.goff = gtotal;
        gtotal += numel;
        param_index[p.name] = (int)params.size();
        *idx = (int)params.size();
        params.push_back(std::move(p));
        return 0;
    }
    float* pf(int idx) { return params[idx].f32; }

    // conv layer with packed bf16 weights [Cout][K0 + C1], K0 = ks*ks*C0; optional shortcut (1x1) segment
    int add_conv(ConvLayer** out, const std::string& name, int C0, int Cout, int ks, int stride, int ups,
                 const std::string& sc_name, int C1a, int C1b) {
        auto L = std::make_unique<ConvLayer>();
        L->name = name; L->C0 = C0; L->Cout = Cout; L->ks = ks; L->stride = stride; L->ups = ups;
        L->C1a = C1a; L->C1b = C1b; L->K0 = ks * ks * C0; L->Ktot = L->K0 + C1a + C1b;
        L->subpixel = ups && use_umma && C0 % 64 == 0 && Cout % 64 == 0;
        if (L->subpixel) { L->K0 = 4 * C0; L->Ktot = 4 * C0; }
        RFV_TRY(dalloc(&L->w, (size_t)Cout * L->Ktot * (L->subpixel ? 4 : 1)));
        RFV_TRY(dalloc(&L->bias, (size_t)Cout));
        int iw, ib, isw = -1, isb = -1;
        RFV_TRY(add_param(name + ".weight", (int64_t)Cout * C0 * ks * ks, &iw));
        RFV_TRY(add_param(name + ".bias", Cout, &ib));
        ConvLayer* l = L.get();
        const int C1 = C1a + C1b;
        if (ks > 1) { params[iw].O = Cout; params[iw].I = C0; params[iw].KK = ks * ks; }
        l->iw = iw; l->ib = ib;
        if (C1 > 0) {
            RFV_TRY(add_param(sc_name + ".weight", (int64_t)Cout * C1, &isw));
            RFV_TRY(add_param(sc_name + ".bias", Cout, &isb));
        }
        l->isw = isw; l->isb = isb;
        l->wparams.push_back(iw);
        if (isw >= 0) l->wparams.push_back(isw);
        params[iw].repack = [this, l, iw](cudaStream_t s) {
            if (l->subpixel) pack_upsample_weight_kernel<<<256, 256, 0, s>>>(pf(iw), l->w, l->Cout, l->C0);
            else pack_conv_weight_kernel<<<256, 256, 0, s>>>(pf(iw), l->w, l->Cout, l->C0, l->ks * l->ks, l->Ktot, 0);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "pack launch failed");
        };
        if (!L->subpixel)  // (the pre-summed sub-pixel taps cannot be un-summed: readback returns the fp32 upload)
            params[iw].readback = [this, l](float* dst, cudaStream_t s) {
                unpack_conv_weight_kernel<<<256, 256, 0, s>>>(l->w, dst, l->Cout, l->C0, l->ks * l->ks, l->Ktot, 0);
                return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "unpack launch failed");
            };
        auto fuse_bias = [this, l, ib, isb](cudaStream_t s) {
            add_vec_kernel<<<(l->Cout + 255) / 256, 256, 0, s>>>(pf(ib), isb >= 0 ? pf(isb) : nullptr, l->bias, l->Cout);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "bias launch failed");
        };
        params[ib].repack = fuse_bias;
        if (C1 > 0) {
            params[isw].repack = [this, l, isw, C1](cudaStream_t s) {
                pack_conv_weight_kernel<<<256, 256, 0, s>>>(pf(isw), l->w, l->Cout, C1, 1, l->Ktot, l->K0);
                return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "pack launch failed");
            };
            params[isw].readback = [this, l, C1](float* dst, cudaStream_t s) {
                unpack_conv_weight_kernel<<<256, 256, 0, s>>>(l->w, dst, l->Cout, C1, 1, l->Ktot, l->K0);
                return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "unpack launch failed");
            };
            params[isb].repack = fuse_bias;
        }
        *out = l;
        convs.push_back(std::move(L));
        return 0;
    }

    // ----- op recording ---------------------------------------------------------------------------------
    void push(const std::string& kind, const std::string& label, double flops,
              std::function<cudaError_t(const RunCtx&, cudaStream_t)> fn) {
        Op op;
        op.kind = kind; op.label = label; op.flops = flops; op.run = std::move(fn);
        if (rec != &ops) {
            op.lane = cur_lane;
            if (cur_lane == 0 && pending_main_sync) { op.sync = 2; pending_main_sync = false; }
            if (cur_lane == 1 && pending_side_sync) { op.sync = 1; pending_side_sync = false; }
        }
        if (rec == &ops) flops_per_image += flops;
        rec->push_back(std::move(op));
    }
    void set_last_bytes(double bytes) { rec->back().bytes = bytes; }

    int make_map4(CUtensorMap* m, const bf16* base, int C, int Wd, int Hd, int Nd, size_t sW, size_t sH, size_t sN,
                  int bw, int bh, int bn) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Nd};
        cuuint64_t strides[3] = {sW * 2, sH * 2, sN * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RFV_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
        return 0;
    }
    int make_map2(CUtensorMap* m, const bf16* base, int K, int rows, int box_rows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RFV_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed with CUresult %d", (int)r);
        return 0;
    }

    // Record one convolution.  in0: segment-0 input; sc: raw shortcut sources (0..2); resid: identity residual.
    // acc_of / acc_k (training, gradient-producing convs): accumulate into `out` in place unless this is the first
    // backward writer of acc_of's gradient, i.e. the last forward consumer (decided at run time: consumer counts are
    // final only once the whole plan is built).
    struct FuseReq { const float* coef; int C; int silu; ActP second; };   // GroupNorm applied to segment 0 inside the conv
    // can this conv take its GroupNorm(+SiLU) inside the kernel?  (weights-as-A kernel, sampling engines only)
    bool can_fuse_gn(int C0, int Cout, int H, int W) const {
        // fuse_mode 1 (default): only where the halo box is <= 1.5x the tile (32-pixel rows: four rows per tile); 2: everywhere
        const bool here = fuse_mode == 2 || (fuse_mode == 1 && W == 32);
        return here && !train && use_umma && use_wa && H == W && (W == 32 || W == 64 || W == 128) && C0 % 64 == 0 && Cout % 64 == 0;
    }
    // Geometry and shared-memory plan of a weights-as-A conv (conv_wa.cuh).  Picks the tile width N (positions per tile) that
    // wastes the fewest MMA columns among those that keep the whole weight matrix resident; layers that have to stream their
    // weights take the widest tiles (weight blocks are re-fetched per tile).
    int plan_wa(WaGeom* gp, size_t* smem, const ConvLayer* L, int W, int H, bool may_resid, int cch0a, bool fuse) {
        WaGeom& g = *gp;
        const bool pair = L->Cout % 128 != 0;
        g.W = W; g.H = H; g.pitch = W + 1;
        g.ctile = pair ? 64 : 128;
        g.n_tiles = L->Cout / g.ctile;
        g.slots0 = pair ? 6 : 9;
        g.cch0 = L->C0 / 64; g.cch0a = cch0a; g.cch1a = L->C1a / 64; g.cch1b = L->C1b / 64;
        g.cchr = 0;   // set per launch
        g.nblk = g.cch0 * g.slots0 + g.cch1a + g.cch1b + g.ctile / 64;
        g.inv_pitch = (uint32_t)((0x100000000ull + g.pitch - 1) / g.pitch);
        const int nblk_max = g.nblk - (may_resid ? 0 : g.ctile / 64);
        const int avail = 227 * 1024 - 2048 - 512 - wa_staging_bytes(pair, fuse);
        const int positions = H * g.pitch;
        double best = 1e30;
        int bestN = 0;
        bool best_res = false;
        auto shape = [&](int N) {
            g.N = N; g.adv = N - (pair ? 1 : 0);
            g.tiles_per_img = (positions + g.adv - 1) / g.adv;
            g.rows = (g.pitch - 1 + N - 1) / g.pitch + 3;
            g.box_bytes = g.rows * g.pitch * 128;
            g.stage_bytes = (g.box_bytes + 128 + 1023) & ~1023;
        };
        for (int N = 256; N >= 128; N -= 32) {
            shape(N);
            const bool res = g.n_tiles == 1 && (long)nblk_max * WA_BLK + 2L * g.stage_bytes <= avail;
            if (!res && 3L * WA_BLK + 2L * g.stage_bytes > avail) continue;
            // streamed weights: every tile re-fetches the whole matrix, so narrow tiles cost L2 bandwidth on top of the columns
            // (measured on the 64->64 layers at N = 160 / 192 / 224: a tile costs its columns plus ~0.9 us of fixed time, about the
            // MMA time of 100 columns)
            const double cost = (double)g.tiles_per_img * (N + 100) / positions * (res ? 1.0 : 1.0 + 0.25 * (256 - N) / 32.0);
            if ((res && !best_res) || (res == best_res && cost < best - 1e-9)) { best = cost; bestN = N; best_res = res; }
        }
        if (!bestN) return fail(RFV_ERR_INVALID, "conv %s: weights-as-A tile does not fit shared memory", L->name.c_str());
        if (const char* ev = getenv("RFV_WA_N")) {   // experiments
            const int Ne = atoi(ev);
            if (Ne >= 128 && Ne <= 256 && Ne % 32 == 0) {
                shape(Ne);
                if (!best_res || (long)nblk_max * WA_BLK + 2L * g.stage_bytes <= avail) bestN = Ne;
            }
        }
        shape(bestN);
        g.inv_tpi = (uint32_t)((0x100000000ull + g.tiles_per_img - 1) / g.tiles_per_img);
        g.tstages = g.N <= 128 ? 4 : (g.N <= 160 ? 3 : 2);
        g.tstride = g.N <= 128 ? 128 : (g.N <= 160 ? 160 : 256);
        if (getenv("RFV_WA_TST")) { g.tstages = 2; g.tstride = 256; }
        g.pf = getenv("RFV_WA_PF") ? atoi(getenv("RFV_WA_PF")) : 0;
        g.dbg = getenv("RFV_WA_DBG") ? atoi(getenv("RFV_WA_DBG")) : 0;
        g.rs = getenv("RFV_WA_RS") ? atoi(getenv("RFV_WA_RS")) : 1;
        if (g.rs < 1 || g.rows % g.rs != 0) g.rs = 1;   // the requests must tile the box exactly (expect_tx = box_bytes)
        g.resident = best_res ? 1 : 0;
        int wregion;
        if (g.resident) { g.w_stages = 1; wregion = nblk_max * WA_BLK; g.a_stages = std::min(4, (avail - wregion) / g.stage_bytes); }
        else {
            g.w_stages = 6;
            while (g.w_stages > 3 && (avail - g.w_stages * WA_BLK) / g.stage_bytes < (g.w_stages > 4 ? 3 : 2)) --g.w_stages;
            if (const char* ev = getenv("RFV_WA_WST")) { if (atoi(ev) >= 2 && (avail - atoi(ev) * WA_BLK) / g.stage_bytes >= 2) g.w_stages = atoi(ev); }
            wregion = g.w_stages * WA_BLK;
            g.a_stages = std::min(4, (avail - wregion) / g.stage_bytes);
        }
        if (g.a_stages < 2) return fail(RFV_ERR_INVALID, "conv %s: weights-as-A tile does not fit shared memory", L->name.c_str());
        g.tblk = 0; g.tcol0 = 0; g.wa = nullptr;
        // Weight blocks in TENSOR MEMORY.  Two accumulator stages of N = 192 columns leave 128 of the 512 columns free: four
        // 128 x 64 weight blocks (32 columns each) move there and are multiplied in TS mode (A operand from TMEM, same rate as
        // the shared-memory form at this N).  That returns 64 KB of shared memory to the box ring, which is what the layers with
        // shortcut / residual chunks lack: with one full-size stage per chunk type the reload of a tile's main box cannot start
        // before its own MMAs have finished (measured: 5,000 cycles per tile against 2,700 of MMAs on the 64->64 + identity
        // layers), and the layers whose blocks do not fit next to two stages otherwise stream 144 KB of weights per tile or fall
        // back to N = 128 tiles.  RFV_WA_TMEM: 0 = off, 1 = where it replaces streamed weights, N <= 160 tiles or a two-stage ring
        // under multi-chunk tiles, 2 (default) = every layer it fits with three box stages (also the single-chunk 64->64 convs).
        // Same-box measurements (tools/ab_tmem.sh, tools/pairs_ab.py): 64->64 + identity 0.266 -> 0.233 ms, 64->64 + 128-channel
        // shortcut 0.265 -> 0.225, 64->128 0.133 -> 0.102 per 512 images; the single-chunk 64->64 convs are 4 % SLOWER timed alone
        // (0.191 -> 0.199: 22 tiles of 192 instead of 17 of 256) but the power-capped production loop is not: 584 / 588 / 590
        // pairs/s for modes 0 / 1 / 2 (TS-mode MMAs read a third less shared memory per instruction).
        static const int tmode = getenv("RFV_WA_TMEM") ? atoi(getenv("RFV_WA_TMEM")) : 2;
        if (tmode > 0 && g.n_tiles == 1 && !getenv("RFV_WA_N")) {
            const WaGeom keep = g;
            const int T = std::min(4, g.slots0);
            const int chunks = g.cch0 + g.cch1a + g.cch1b + (may_resid ? g.ctile / 64 : 0);
            shape(192);
            const long wbytes = (long)(nblk_max - T) * WA_BLK;
            const int st = wbytes >= 0 ? (int)std::min<long>(4, (avail - wbytes) / g.stage_bytes) : 0;
            bool take = false;
            // three stages or nothing: with two, a 64->64 conv with a 192-channel shortcut (four chunks per tile) measured 0.358 ms
            // against 0.344 ms on streamed weights at N = 256
            if (st >= 3) take = !keep.resident || keep.N <= 160 || (chunks >= 2 && keep.a_stages < 3) || tmode >= 2;
            if (take) {
                g.inv_tpi = (uint32_t)((0x100000000ull + g.tiles_per_img - 1) / g.tiles_per_img);
                g.tstages = 2; g.tstride = 192;
                g.tblk = T; g.tcol0 = 384;
                g.resident = 1; g.w_stages = 1; g.a_stages = st;
                if (g.rows % g.rs != 0) g.rs = 1;
                wregion = (int)wbytes;
            } else {
                g = keep;
            }
        }
        *smem = 2048 + (size_t)g.a_stages * g.stage_bytes + wregion + wa_staging_bytes(pair, fuse) + 512;
        return 0;
    }
    // allocate the block-layout weight copy of a layer and chain its refresh behind the repack of the parameters it derives from
    int ensure_wa(ConvLayer* L, const WaGeom& g) {
        if (L->wa) return 0;
        RFV_TRY(dalloc(&L->wa, (size_t)g.n_tiles * g.nblk * (WA_BLK / 2)));
        const int pair = g.ctile == 64 ? 1 : 0, cch1 = g.cch1a + g.cch1b, cchr = g.ctile / 64, cch0 = g.cch0, n_tiles = g.n_tiles;
        for (int pi : L->wparams) {
            auto prev = params[pi].repack;
            params[pi].repack = [this, prev, L, pair, cch0, cch1, cchr, n_tiles](cudaStream_t s) {
                const int rc = prev ? prev(s) : 0;
                if (rc) return rc;
                pack_wa_kernel<<<256, 256, 0, s>>>(L->w, L->wa, L->Cout, L->C0, L->K0, L->Ktot, cch0, cch1, cchr, pair, n_tiles);
                return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "weights-as-A pack launch failed");
            };
        }
        return 0;
    }

    int conv_op(ConvLayer* L, ActP in0, std::vector<ActP> sc, ActP resid, ActP out, int temb_off, bool want_stats,
                ActP acc_of = nullptr, int acc_k = 0, const FuseReq* fr = nullptr) {
        const std::string pre = rec == &ops ? "conv:" : "bwd:dgrad:";
        ConvParams p{};
        p.out = out->p; p.a0 = in0->p;
        p.s1a = sc.size() > 0 ? sc[0]->p : nullptr;
        p.s1b = sc.size() > 1 ? sc[1]->p : nullptr;
        p.w = L->w; p.bias = L->bias;
        p.temb = temb_off >= 0 ? tproj + temb_off : nullptr;
        p.resid = resid ? resid->p : nullptr;
        p.stats = want_stats ? out->stats : nullptr;
        p.Ho = out->H; p.Wo = out->W; p.Cout = L->Cout;
        p.H0 = in0->H; p.W0 = in0->W; p.C0 = L->C0;
        p.ks = L->ks; p.stride = L->stride; p.ups = L->ups;
        p.C1a = L->C1a; p.C1b = L->C1b; p.K0 = L->K0; p.Ktot = L->Ktot;
        p.slab_shift = slab_shift;
        const double fl = 2.0 * ((double)L->ks * L->ks * L->C0 + L->C1a + L->C1b) * L->Cout * out->H * out->W;  // algorithmic
        const int HoWo = out->H * out->W;
        if (HoWo % 32 != 0) return fail(RFV_ERR_INVALID, "conv %s: Ho*Wo=%d must be a multiple of 32", L->name.c_str(), HoWo);

        // tile geometry lives on the grid the GEMM rows enumerate: the output pixels, or the INPUT pixels of an
        // upsample conv (each input-resolution box yields four output phases)
        const int gW = L->ups ? in0->W : out->W, gH = L->ups ? in0->H : out->H;
        const bool pow2 = is_pow2(gW) && is_pow2(gH);
        const bool umma_ok = use_umma && pow2 && (!L->ups || L->subpixel) && L->C0 % 64 == 0 && L->C1a % 64 == 0 &&
                             L->C1b % 64 == 0 && L->Cout % 64 == 0 && gH * gW >= 64 && gW >= 8 && (L->stride == 1 || sc.empty());
        if (L->subpixel && !umma_ok) return fail(RFV_ERR_INVALID, "conv %s: sub-pixel packing needs the tcgen05 path", L->name.c_str());
        const bool wa_ok = umma_ok && use_wa && L->ks == 3 && L->stride == 1 && !L->ups && out->W == out->H &&
                           (out->W == 32 || out->W == 64 || out->W == 128) && (!resid || resid->C == L->Cout);
        if (wa_ok) {
            // weights-as-A kernel (conv_wa.cuh): N = up to 256 positions per MMA, optional in-kernel GroupNorm on segment 0
            struct WBundle { CUtensorMap a0, a0b, a1, a2, r, w; WaGeom g; size_t smem; bool pair, fuse; const void* wa; };
            auto bd = std::make_shared<WBundle>();
            WaGeom& g = bd->g;
            bd->pair = L->Cout % 128 != 0;
            bd->fuse = fr != nullptr;
            RFV_TRY(plan_wa(&g, &bd->smem, L, out->W, out->H, resid != nullptr || acc_of != nullptr, in0->C / 64, fr != nullptr));
            RFV_TRY(ensure_wa(L, g));
            bd->wa = L->wa;
            auto amap = [&](CUtensorMap* m, const ActP& t) {
                return make_map4(m, t->p, t->C, t->W, t->H, cap, t->C, (size_t)t->W * t->C, (size_t)t->H * t->W * t->C, g.pitch, g.rs, 1);
            };
            RFV_TRY(amap(&bd->a0, in0));
            bd->a0b = bd->a0; bd->a1 = bd->a0; bd->a2 = bd->a0;
            if (fr && fr->second) RFV_TRY(amap(&bd->a0b, fr->second));
            if (sc.size() > 0) RFV_TRY(amap(&bd->a1, sc[0]));
            if (sc.size() > 1) RFV_TRY(amap(&bd->a2, sc[1]));
            RFV_TRY(amap(&bd->r, resid ? resid : out));   // identity residual, or in-place gradient accumulation (acc_of)
            RFV_TRY(make_map2(&bd->w, L->wa, 64, g.n_tiles * g.nblk * 128, 128));
            if (fr) { p.gn_coef = fr->coef; p.gn_C = fr->C; p.gn_silu = fr->silu; }
            const int sms = num_sms, sumC_ = sumC;
            push("conv_wa", pre + L->name, fl, [p, bd, sms, sumC_, acc_of, acc_k](const RunCtx& rc, cudaStream_t s) mutable {
                ConvParams q = p;
                q.B = rc.B;
                if (acc_of && acc_k != acc_of->consumers - 1) q.resid = q.out;
                q.temb_stride = rc.t ? sumC_ : 0;
                if (rc.temb_row > 0 && q.temb) q.temb += (size_t)rc.temb_row * sumC_;
                WaGeom g = bd->g;
                g.cchr = q.resid ? g.ctile / 64 : 0;
                if (g.cchr && !getenv("RFV_WA_PF")) g.pf = 1;   // two boxes per tile through a short ring: prefetch the next tile's into L2
                g.m_tiles = rc.B * g.tiles_per_img;
                g.wa = bd->wa;
                const int grid = std::min(g.m_tiles * g.n_tiles, sms);
                auto kern = bd->fuse ? (bd->pair ? conv_wa_kernel<true, true> : conv_wa_kernel<false, true>)
                                     : (bd->pair ? conv_wa_kernel<true, false> : conv_wa_kernel<false, false>);
                return klaunch(rc.pdl, kern, dim3(grid), dim3(wa_threads(bd->fuse)), bd->smem, s, bd->a0, bd->a0b, bd->a1, bd->a2, bd->r, bd->w, q, g);
            });
        } else if (fr) {
            return fail(RFV_ERR_STATE, "internal: conv %s cannot fuse its GroupNorm", L->name.c_str());
        } else if (umma_ok) {
            struct Bundle { CUtensorMap a0, a1, a2, a3, w; UmmaGeom g; int BN; int max_clusters; bool pair; int pair_clusters; };
            auto bd = std::make_shared<Bundle>();
            UmmaGeom& g = bd->g;
            const int bw = std::min(gW, 128), bh = std::min(gH, 128 / bw), bn = 128 / (bw * bh);
            g.bw_shift = ilog2(bw); g.bh_shift = ilog2(bh);
            g.tiles_w = gW / bw; g.tiles_h = gH / bh;
            g.cch0 = L->C0 / 64; g.cch1a = L->C1a / 64; g.cch1b = L->C1b / 64;
            g.taps = L->subpixel ? 4 : L->ks * L->ks; g.stride2 = (L->stride == 2); g.ups = L->subpixel ? 1 : 0;
            const int BN = (L->Cout % 256 == 0) ? 256 : (L->Cout % 128 == 0 ? 128 : 64);
            bd->BN = BN;
            g.n_tiles = L->Cout / BN;
            const int capN = cap;
            if (!g.stride2) {
                RFV_TRY(make_map4(&bd->a0, in0->p, in0->C, in0->W, in0->H, capN, in0->C, (size_t)in0->W * in0->C,
                                  (size_t)in0->H * in0->W * in0->C, bw, bh, bn));
                bd->a1 = bd->a0; bd->a2 = bd->a0; bd->a3 = bd->a0;
                if (sc.size() > 0)
                    RFV_TRY(make_map4(&bd->a1, sc[0]->p, sc[0]->C, sc[0]->W, sc[0]->H, capN, sc[0]->C, (size_t)sc[0]->W * sc[0]->C,
                                      (size_t)sc[0]->H * sc[0]->W * sc[0]->C, bw, bh, bn));
                if (sc.size() > 1)
                    RFV_TRY(make_map4(&bd->a2, sc[1]->p, sc[1]->C, sc[1]->W, sc[1]->H, capN, sc[1]->C, (size_t)sc[1]->W * sc[1]->C,
                                      (size_t)sc[1]->H * sc[1]->W * sc[1]->C, bw, bh, bn));
            } else {
                // parity views of the (even-sized) input: view(ph,pw)[n,i,j,c] = in[n, 2i+ph, 2j+pw, c]
                const int C = in0->C, Wi = in0->W, Hi = in0->H;
                CUtensorMap* mp[4] = {&bd->a0, &bd->a1, &bd->a2, &bd->a3};
                for (int ph = 0; ph < 2; ++ph)
                    for (int pw = 0; pw < 2; ++pw)
                        RFV_TRY(make_map4(mp[ph * 2 + pw], in0->p + ((size_t)ph * Wi + pw) * C, C, Wi / 2, Hi / 2, capN,
                                          (size_t)2 * C, (size_t)2 * Wi * C, (size_t)Hi * Wi * C, bw, bh, bn));
            }
            g.cluster = cluster;
            // 256-output-channel stride-1 convs run on CTA pairs (conv_umma2.cuh: one M = 256 tcgen05.mma per pair, each CTA
            // stages half of the weight slice)
            // (not the 1x1 convs: four K chunks per tile leave them epilogue-bound, and the pair's hand-shakes cost 2-3 % there)
            bd->pair = (BN == 256 || BN == 128) && !g.stride2 && (g.taps == 9 || g.taps == 4) && cluster == 1 && !(cfg.flags & RFV_FLAG_NO_CTA_PAIR);
            bd->pair_clusters = 0;
            if (bd->pair) {
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3(num_sms / 2 * 2);
                lc.blockDim = dim3(UMMA_THREADS);
                lc.dynamicSmemBytes = BN == 256 ? PairCfg<256>::SMEM_BYTES : PairCfg<128>::SMEM_BYTES;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                int nc = 0;
                const cudaError_t ce = BN == 256 ? cudaOccupancyMaxActiveClusters(&nc, conv_umma2_kernel<256>, &lc)
                                                 : cudaOccupancyMaxActiveClusters(&nc, conv_umma2_kernel<128>, &lc);
                if (ce != cudaSuccess || nc < 1) { cudaGetLastError(); bd->pair = false; }
                else bd->pair_clusters = std::min(nc, num_sms / 2);
            }
            RFV_TRY(make_map2(&bd->w, L->w, L->Ktot, L->Cout * (L->subpixel ? 4 : 1), bd->pair ? BN / 2 : BN / g.cluster));
            bd->max_clusters = num_sms / g.cluster;
            if (g.cluster > 1) {  // how many clusters of this kernel can be co-resident (GPC boundaries strand SMs)
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3(num_sms / g.cluster * g.cluster);
                lc.blockDim = dim3(UMMA_THREADS);
                lc.dynamicSmemBytes = BN == 256 ? UmmaCfg<256>::SMEM_BYTES : (BN == 128 ? UmmaCfg<128>::SMEM_BYTES : UmmaCfg<64>::SMEM_BYTES);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = g.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                int nc = 0;
                cudaError_t ce = BN == 256 ? cudaOccupancyMaxActiveClusters(&nc, conv_umma_kernel<256>, &lc)
                               : BN == 128 ? cudaOccupancyMaxActiveClusters(&nc, conv_umma_kernel<128>, &lc)
                                           : cudaOccupancyMaxActiveClusters(&nc, conv_umma_kernel<64>, &lc);
                if (ce != cudaSuccess || nc < 1) return fail(RFV_ERR_CUDA, "cudaOccupancyMaxActiveClusters failed: %s", cudaGetErrorString(ce));
                bd->max_clusters = std::min(nc, num_sms / g.cluster);
            }
            const int sms = num_sms;
            const int sumC_ = sumC;
            const int gHW = gH * gW;
            push("conv_umma", pre + L->name, fl, [p, bd, gHW, sms, sumC_, acc_of, acc_k](const RunCtx& rc, cudaStream_t s) mutable {
                ConvParams q = p;
                q.B = rc.B;
                if (acc_of && acc_k != acc_of->consumers - 1) q.resid = q.out;
                q.temb_stride = rc.t ? sumC_ : 0;
                if (rc.temb_row > 0 && q.temb) q.temb += (size_t)rc.temb_row * sumC_;
                UmmaGeom g = bd->g;
                g.m_tiles = (int)(((size_t)rc.B * gHW + 127) / 128);
                if (bd->pair) {
                    const int pairs = ((g.m_tiles + 1) / 2) * g.n_tiles * (g.ups ? 4 : 1);
                    cudaLaunchConfig_t lp{};
                    lp.gridDim = dim3(std::min(pairs, bd->pair_clusters) * 2);
                    lp.blockDim = dim3(UMMA_THREADS);
                    lp.stream = s;
                    cudaLaunchAttribute ap[2];
                    ap[0].id = cudaLaunchAttributeClusterDimension;
                    ap[0].val.clusterDim.x = 2; ap[0].val.clusterDim.y = 1; ap[0].val.clusterDim.z = 1;
                    ap[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    ap[1].val.programmaticStreamSerializationAllowed = 1;
                    lp.attrs = ap;
                    lp.numAttrs = rc.pdl ? 2 : 1;
                    if (bd->BN == 256) {
                        lp.dynamicSmemBytes = PairCfg<256>::SMEM_BYTES;
                        return cudaLaunchKernelEx(&lp, conv_umma2_kernel<256>, bd->a0, bd->a1, bd->a2, bd->w, q, g);
                    }
                    lp.dynamicSmemBytes = PairCfg<128>::SMEM_BYTES;
                    return cudaLaunchKernelEx(&lp, conv_umma2_kernel<128>, bd->a0, bd->a1, bd->a2, bd->w, q, g);
                }
                const int super_tiles = ((g.m_tiles + g.cluster - 1) / g.cluster) * g.n_tiles * (g.ups ? 4 : 1);
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3(std::min(super_tiles, bd->max_clusters) * g.cluster);
                lc.blockDim = dim3(UMMA_THREADS);
                lc.stream = s;
                cudaLaunchAttribute at[2];
                int na = 0;
                if (g.cluster > 1) {
                    at[na].id = cudaLaunchAttributeClusterDimension;
                    at[na].val.clusterDim.x = g.cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
                    ++na;
                }
                if (rc.pdl) {
                    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[na].val.programmaticStreamSerializationAllowed = 1;
                    ++na;
                }
                lc.attrs = at;
                lc.numAttrs = na;
                switch (bd->BN) {
                    case 256:
                        lc.dynamicSmemBytes = UmmaCfg<256>::SMEM_BYTES;
                        return cudaLaunchKernelEx(&lc, conv_umma_kernel<256>, bd->a0, bd->a1, bd->a2, bd->a3, bd->w, q, g);
                    case 128:
                        lc.dynamicSmemBytes = UmmaCfg<128>::SMEM_BYTES;
                        return cudaLaunchKernelEx(&lc, conv_umma_kernel<128>, bd->a0, bd->a1, bd->a2, bd->a3, bd->w, q, g);
                    default:
                        lc.dynamicSmemBytes = UmmaCfg<64>::SMEM_BYTES;
                        return cudaLaunchKernelEx(&lc, conv_umma_kernel<64>, bd->a0, bd->a1, bd->a2, bd->a3, bd->w, q, g);
                }
            });
        } else {
            if (L->C0 % 8 != 0 || L->C1a % 8 != 0 || L->C1b % 8 != 0 || L->Cout % 64 != 0)
                return fail(RFV_ERR_INVALID, "conv %s: unsupported channel counts", L->name.c_str());
            const int sumC_ = sumC;
            const int cout = L->Cout;
            push("conv_mma", pre + L->name, fl, [p, HoWo, sumC_, cout, acc_of, acc_k](const RunCtx& rc, cudaStream_t s) {
                ConvParams q = p;
                q.B = rc.B;
                if (acc_of && acc_k != acc_of->consumers - 1) q.resid = q.out;
                q.temb_stride = rc.t ? sumC_ : 0;
                if (rc.temb_row > 0 && q.temb) q.temb += (size_t)rc.temb_row * sumC_;
                dim3 grid((unsigned)(((size_t)rc.B * HoWo + MMA_BM - 1) / MMA_BM), cout / MMA_BN);
                return klaunch(rc.pdl, conv_mma_kernel, grid, dim3(256), 0, s, q);
            });
        }
        return 0;
    }

    struct NormSite {  // what the GroupNorm backward needs to know about one forward norm
        std::vector<ActP> srcs;
        int ig = -1, ib = -1, C = 0, HW = 0, id = 0;
        bool silu = false, drop = false;
        float* cs = nullptr;  // [cap][C][2]
    };
    int gn_op(const std::string& name, std::vector<ActP> srcs, ActP out, bool silu, bool drop = false, NormSite* site = nullptr) {
        int ig, ib;
        const int C = out->C;
        RFV_TRY(add_param(name + ".weight", C, &ig));
        RFV_TRY(add_param(name + ".bias", C, &ib));
        const int site_id = ++norm_sites;
        if (site) {
            site->srcs = srcs; site->ig = ig; site->ib = ib; site->C = C; site->HW = out->H * out->W; site->id = site_id;
            site->silu = silu; site->drop = drop;
            if (train) {
                const size_t n = (size_t)cap * C * 2 + cap + 4;   // sums + per-image arrival counters + ticket (gn_bwd_fused_kernel)
                if (cs_used + n > cs_floats) return fail(RFV_ERR_NOMEM, "GroupNorm-backward arena exhausted");
                site->cs = cs_arena + cs_used;
                cs_used += n;
            }
        }
        const bf16* xa = srcs[0]->p;
        const bf16* xb = srcs.size() > 1 ? srcs[1]->p : nullptr;
        const float* sa = srcs[0]->stats;
        const float* sb = srcs.size() > 1 ? srcs[1]->stats : nullptr;
        if (!sa || (xb && !sb)) return fail(RFV_ERR_STATE, "gn %s: source without statistics", name.c_str());
        const int Ca = srcs[0]->C, Cb = srcs.size() > 1 ? srcs[1]->C : 0;
        const int HW = out->H * out->W;
        bf16* o = out->p;
        const float *gam = pf(ig), *bet = pf(ib);
        const int ss = slab_shift;
        if ((C / 8) % (1 << ss) != 0 || Ca % (1 << ss) != 0) return fail(RFV_ERR_INVALID, "gn %s: group/slab mismatch", name.c_str());
        // ~128 KB of bf16 per block (amortises the per-block scale/shift prologue); threads = multiple of C/8
        int ppb = std::max(1, 65536 / C);
        ppb = std::min(ppb, HW);
        const int vpp = C / 8;
        if (vpp > 256) return fail(RFV_ERR_INVALID, "gn %s: more than 2048 channels unsupported", name.c_str());
        const int threads = (256 / vpp) * vpp;
        const int sms_ = num_sms;
        const int silu_mode = (cfg.flags & RFV_FLAG_SILU_EXP) ? 1 : 2;
        push("gn_apply", "gn:" + name, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
            int ppb_run = ppb;   // small batches: smaller pixel slabs so that the grid still fills the GPU (>= 8 blocks per SM)
            while (ppb_run > 64 && ((HW + ppb_run - 1) / ppb_run) * rc.B < 8 * sms_) ppb_run /= 2;
            dim3 grid((HW + ppb_run - 1) / ppb_run, rc.B);
            const uint32_t dt_ = drop ? rc.drop_thresh : 0u;
            return klaunch(rc.pdl, gn_apply_kernel, grid, dim3(threads), 2 * C * sizeof(float), s, xa, xb, sa, sb, gam, bet, o, Ca, Cb, HW, ss,
                           silu ? silu_mode : 0, ppb_run, 1e-5f, dt_, rc.seed ^ ((uint32_t)site_id * 0x9E3779B9u), rc.drop_scale);
        });
        set_last_bytes(4.0 * C * HW);   // 2 B read + 2 B written per element
        return 0;
    }

    // GroupNorm site whose apply pass is fused into the consuming conv: registers the affine parameters (same order as gn_op)
    // and records the tiny per-(image, channel) coefficient kernel.
    int gn_coef_op(const std::string& name, std::vector<ActP> srcs, int HW, float** coef_out) {
        int ig, ib, C = 0;
        for (auto& a : srcs) C += a->C;
        RFV_TRY(add_param(name + ".weight", C, &ig));
        RFV_TRY(add_param(name + ".bias", C, &ib));
        ++norm_sites;
        float* coef = nullptr;
        RFV_TRY(dalloc(&coef, (size_t)cap * C * 2));
        const float* sa = srcs[0]->stats;
        const float* sb = srcs.size() > 1 ? srcs[1]->stats : nullptr;
        if (!sa || (srcs.size() > 1 && !sb)) return fail(RFV_ERR_STATE, "gn %s: source without statistics", name.c_str());
        const int Ca = srcs[0]->C, Cb = srcs.size() > 1 ? srcs[1]->C : 0, ss = slab_shift;
        const float *gam = pf(ig), *bet = pf(ib);
        push("gn_coef", "gn:" + name, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
            return klaunch(rc.pdl, gn_coef_kernel, dim3(rc.B), dim3(256), 0, s, sa, sb, gam, bet, coef, Ca, Cb, HW, ss, 1e-5f);
        });
        *coef_out = coef;
        return 0;
    }

    // ResidualBlock (models/unet.py:55-64).  srcs: 1 tensor, or 2 for the decoder's virtual concat [h, skip].
    int res_block(const std::string& name, std::vector<ActP> srcs, int Cout, int* temb_cursor, ActP* result) {
        int Cin = 0;
        for (auto& a : srcs) Cin += a->C;
        const int H = srcs[0]->H, W = srcs[0]->W;
        ActP a1, h, a2, out;
        NormSite n1, n2;
        const int ka = srcs[0]->consumers++, kb = srcs.size() > 1 ? srcs[1]->consumers++ : 0;
        const bool fuse1 = can_fuse_gn(Cin, Cout, H, W), fuse2 = can_fuse_gn(Cout, Cout, H, W);
        float* coef1 = nullptr;
        if (fuse1) RFV_TRY(gn_coef_op(name + ".norm1", srcs, H * W, &coef1));
        else {
            RFV_TRY(new_act(&a1, Cin, H, W, false));
            RFV_TRY(gn_op(name + ".norm1", srcs, a1, true, false, &n1));
        }
        ConvLayer *c1, *c2;
        RFV_TRY(add_conv(&c1, name + ".conv1", Cin, Cout, 3, 1, 0, "", 0, 0));
        RFV_TRY(new_act(&h, Cout, H, W, true));
        // per-block time projection rows live at [temb_cursor, temb_cursor + Cout) of the concatenated matrix
        int iw, ib;
        RFV_TRY(add_param(name + ".time_mlp.1.weight", (int64_t)Cout * td, &iw));
        RFV_TRY(add_param(name + ".time_mlp.1.bias", Cout, &ib));
        const int off = *temb_cursor;
        *temb_cursor += Cout;
        time_projs.push_back({off, Cout, iw, ib, param_index["velocity_net." + name + ".conv1.bias"]});
        params[iw].repack = [this, iw, off, Cout](cudaStream_t s) {
            cudaError_t e = cudaMemcpyAsync(wcat + (size_t)off * td, pf(iw), (size_t)Cout * td * sizeof(float), cudaMemcpyDeviceToDevice, s);
            return e == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "wcat copy failed");
        };
        // bcat rows = time_mlp bias + conv1 bias: the conv epilogue then adds ONE vector per channel
        const int icb = param_index["velocity_net." + name + ".conv1.bias"];
        auto fuse = [this, ib, icb, off, Cout](cudaStream_t s) {
            add_vec_kernel<<<(Cout + 255) / 256, 256, 0, s>>>(pf(ib), pf(icb), bcat + off, Cout);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "bcat fuse failed");
        };
        params[ib].repack = fuse;
        {
            auto prev = params[icb].repack;
            params[icb].repack = [prev, fuse](cudaStream_t s) { int rc = prev ? prev(s) : 0; return rc ? rc : fuse(s); };
        }
        if (fuse1) {
            FuseReq fr{coef1, Cin, 1, srcs.size() > 1 ? srcs[1] : nullptr};
            RFV_TRY(conv_op(c1, srcs[0], {}, nullptr, h, off, true, nullptr, 0, &fr));
        } else {
            RFV_TRY(conv_op(c1, a1, {}, nullptr, h, off, true));
            release(a1);
        }
        float* coef2 = nullptr;
        if (fuse2) RFV_TRY(gn_coef_op(name + ".norm2", {h}, H * W, &coef2));
        else {
            RFV_TRY(new_act(&a2, Cout, H, W, false));
            RFV_TRY(gn_op(name + ".norm2", {h}, a2, true, true, &n2));  // nn.Dropout follows this SiLU (models/unet.py:62)
        }
        if (!fuse2) release(h);
        RFV_TRY(new_act(&out, Cout, H, W, true));
        if (Cin != Cout) {
            RFV_TRY(add_conv(&c2, name + ".conv2", Cout, Cout, 3, 1, 0, name + ".shortcut", srcs[0]->C, srcs.size() > 1 ? srcs[1]->C : 0));
            if (fuse2) { FuseReq fr{coef2, Cout, 1, nullptr}; RFV_TRY(conv_op(c2, h, srcs, nullptr, out, -1, true, nullptr, 0, &fr)); }
            else RFV_TRY(conv_op(c2, a2, srcs, nullptr, out, -1, true));
        } else {
            RFV_TRY(add_conv(&c2, name + ".conv2", Cout, Cout, 3, 1, 0, "", 0, 0));
            if (fuse2) { FuseReq fr{coef2, Cout, 1, nullptr}; RFV_TRY(conv_op(c2, h, {}, srcs[0], out, -1, true, nullptr, 0, &fr)); }
            else RFV_TRY(conv_op(c2, a2, {}, srcs[0], out, -1, true));
        }
        if (fuse2) release(h);
        else release(a2);
        named_acts[name] = out;
        *result = out;
        if (train) {
            // ---- backward of the block (models/unet.py:55-64 reversed); out->grad is complete when this runs ----
            RFV_TRY(ensure_grad(out));
            begin_bwd();
            ActP dOut = grad_view(out), Tsc, Ta2, Th, Ta1;
            ConvLayer *g1, *g2, *gs = nullptr;
            RFV_TRY(add_dgrad_layer(&g2, c2, 0));
            RFV_TRY(add_dgrad_layer(&g1, c1, 0));
            if (Cin != Cout) {
                RFV_TRY(add_dgrad_layer(&gs, c2, 1));
                RFV_TRY(scratch_act(&Tsc, 0, Cin, H, W));
                RFV_TRY(conv_op(gs, dOut, {}, nullptr, Tsc, -1, false));
                on_side(false);
                RFV_TRY(bwd_wgrad(name + ".shortcut", 1, srcs[0], out->grad, Cout, W, H, c2->isw, Cin, 0));
                if (srcs.size() > 1) RFV_TRY(bwd_wgrad(name + ".shortcut/b", 1, srcs[1], out->grad, Cout, W, H, c2->isw, Cin, srcs[0]->C));
            }
            on_side(false);
            RFV_TRY(bwd_wgrad(name + ".conv2", 0, a2, out->grad, Cout, W, H, c2->iw, 9 * Cout, 0));
            bwd_colsum(name + ".conv2.bias", out->grad, Cout, H * W, nullptr, 0, c2->ib, c2->isb);
            on_main();
            RFV_TRY(scratch_act(&Ta2, 1, Cout, H, W));
            RFV_TRY(conv_op(g2, dOut, {}, nullptr, Ta2, -1, false));
            RFV_TRY(scratch_act(&Th, 2, Cout, H, W));
            // also d(time projection)[n][c] = sum over pixels of dh (conv1.bias and time_mlp.1.bias get its batch sum later)
            RFV_TRY(bwd_gn(name + ".norm2", n2, Ta2->p, nullptr, nullptr, 0, 0, Th->p, d_tproj + off, sumC));
            on_side(true);    // reads dh, which main has just produced
            RFV_TRY(bwd_wgrad(name + ".conv1", 0, a1, Th->p, Cout, W, H, c1->iw, 9 * Cin, 0));
            on_main();
            RFV_TRY(scratch_act(&Ta1, 1, Cin, H, W));
            RFV_TRY(conv_op(g1, Th, {}, nullptr, Ta1, -1, false));
            RFV_TRY(bwd_gn(name + ".norm1", n1, Ta1->p, Tsc ? Tsc->p : nullptr, Cin == Cout ? out->grad : nullptr, ka, kb));
            end_bwd();
        }
        return 0;
    }


    // ===== training: backward recording (each forward stage appends one block; blocks run back to front) =====
    // Two streams in the backward pass: data gradients and GroupNorm backward on the main one, weight / bias gradients on a
    // side one (the tcgen05 wgrad kernel leaves the issue slots and HBM idle, the GroupNorm backward leaves the tensor pipe
    // idle, and both fit on an SM together).  The first main op of a stage waits for all earlier side work (the scratch
    // tensors it overwrites may still be read there); a side op waits for main where it consumes what main just produced.
    int cur_lane = 0;
    bool pending_main_sync = false, pending_side_sync = false;
    void begin_bwd() { bwd_blocks.emplace_back(); rec = &bwd_blocks.back(); cur_lane = 0; pending_main_sync = true; pending_side_sync = true; }
    void end_bwd() { rec = &ops; cur_lane = 0; }
    void on_side(bool wait_main) { cur_lane = 1; if (wait_main) pending_side_sync = true; }
    void on_main() { cur_lane = 0; }
    float* gslot(int pi) { return gflat + params[pi].goff; }  // run time only (the flat buffer is allocated after the plan)
    // record time: the op being recorded writes parameter pi's gradient; the LAST block to do so (lowest index: blocks run back
    // to front) decides when the slot is final -- the order the slots are laid out in and the gradient buckets are cut by
    void mark_grad(int pi) {
        if (pi >= 0 && !bwd_blocks.empty()) params[pi].final_block = std::min(params[pi].final_block, (int)bwd_blocks.size() - 1);
    }

    int ensure_grad(const ActP& a) {
        if (a->grad) return 0;
        return dalloc(&a->grad, a->bytes / sizeof(bf16));
    }
    ActP view(bf16* ptr, int C, int H, int W) {
        auto a = std::make_shared<Act>();
        a->p = ptr; a->C = C; a->H = H; a->W = W; a->refs = 1;
        a->bytes = (size_t)cap * H * W * C * sizeof(bf16);
        return a;
    }
    int scratch_act(ActP* out, int idx, int C, int H, int W) {
        if ((size_t)C * H * W > scratch_elems) return fail(RFV_ERR_STATE, "internal: backward scratch too small for %dx%dx%d", C, H, W);
        *out = view(scratch[idx], C, H, W);
        return 0;
    }
    ActP grad_view(const ActP& a) { return view(a->grad, a->C, a->H, a->W); }

    // Data-gradient twin of a forward conv: the same kernels on repacked weights.
    // kind 0: segment 0 transposed + spatially flipped (3x3 or 1x1, stride 1); 1: the fused 1x1 shortcut segment
    // transposed; 2: stride-2 conv as four sub-pixel phases over dY; 3: nearest-x2-upsample + 3x3 conv as a 16-tap
    // stride-2 conv over the four parity views of dY (the adjoint of the sub-pixel forward).
    int add_dgrad_layer(ConvLayer** out, ConvLayer* src, int kind) {
        auto L = std::make_unique<ConvLayer>();
        const int C1 = src->C1a + src->C1b;
        L->name = src->name + (kind == 1 ? "+shortcut" : "");
        L->stride = 1;
        if (kind == 0) { L->C0 = src->Cout; L->Cout = src->C0; L->ks = src->ks; L->K0 = L->ks * L->ks * L->C0; }
        else if (kind == 1) { L->C0 = src->Cout; L->Cout = C1; L->ks = 1; L->K0 = L->C0; }
        else if (kind == 2) { L->C0 = src->Cout; L->Cout = src->C0; L->ks = 3; L->ups = 1; L->subpixel = true; L->K0 = 4 * L->C0; }
        else { L->C0 = src->Cout; L->Cout = src->C0; L->ks = 4; L->stride = 2; L->K0 = 16 * L->C0; }  // 3: upsample conv
        L->Ktot = L->K0;
        RFV_TRY(dalloc(&L->w, (size_t)L->Cout * L->Ktot * (L->subpixel ? 4 : 1)));
        L->bias = zero_bias;
        ConvLayer* l = L.get();
        const int pi = kind == 1 ? src->isw : src->iw;
        l->wparams.push_back(pi);
        auto prev = params[pi].repack;
        params[pi].repack = [this, prev, l, src, kind, pi, C1](cudaStream_t s) {
            const int rc = prev ? prev(s) : 0;
            if (rc) return rc;
            if (kind == 0) pack_conv_weight_T_kernel<<<256, 256, 0, s>>>(pf(pi), l->w, src->Cout, src->C0, src->ks * src->ks, 0, 0);
            else if (kind == 1) pack_conv_weight_T_kernel<<<256, 256, 0, s>>>(pf(pi), l->w, src->Cout, C1, 1, 0, 0);
            else if (kind == 2) pack_down_dgrad_weight_kernel<<<256, 256, 0, s>>>(pf(pi), l->w, src->Cout, src->C0);
            else pack_up_dgrad_weight_kernel<<<256, 256, 0, s>>>(pf(pi), l->w, src->Cout, src->C0);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "dgrad weight pack launch failed");
        };
        *out = l;
        convs.push_back(std::move(L));
        return 0;
    }

    // dW slot of parameter `pi` (+)= wgrad(dy, a).  kind as make_wgrad_geom; a: the conv's input activations (for
    // kind 2 the full-resolution tensor); dy: [cap, Ho, Wo, Cout].  The slot row holds ldw floats, this conv's columns
    // start at koff_base (second source of a shortcut over a virtual concat).
    int bwd_wgrad(const std::string& label, int kind, const ActP& a, const bf16* dy, int Cout, int Wo, int Ho, int pi, int ldw, int koff_base,
                  float* dst_override = nullptr) {
        struct Bundle { CUtensorMap ma[4], my[4]; WgradGeom g; size_t smem; };
        auto bd = std::make_shared<Bundle>();
        if (a->C % 64 != 0 || Cout % 64 != 0) return fail(RFV_ERR_INVALID, "wgrad %s: channel counts must be multiples of 64", label.c_str());
        if (!make_wgrad_geom(&bd->g, Wo, Ho, a->C, Cout, kind, ldw, koff_base)) return fail(RFV_ERR_INVALID, "wgrad %s: tile does not fit shared memory", label.c_str());
        const WgradGeom& g = bd->g;
        if (kind != 2) {
            RFV_TRY(make_map4(&bd->ma[0], a->p, a->C, a->W, a->H, cap, a->C, (size_t)a->W * a->C, (size_t)a->H * a->W * a->C, g.pitch, g.R + 2, 1));
            bd->ma[1] = bd->ma[2] = bd->ma[3] = bd->ma[0];
        } else {
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw)
                    RFV_TRY(make_map4(&bd->ma[ph * 2 + pw], a->p + ((size_t)ph * a->W + pw) * a->C, a->C, Wo, Ho, cap, (size_t)2 * a->C,
                                      (size_t)2 * a->W * a->C, (size_t)a->H * a->W * a->C, g.pitch, g.R + 2, 1));
        }
        if (kind != 3) {
            RFV_TRY(make_map4(&bd->my[0], dy, Cout, Wo, Ho, cap, Cout, (size_t)Wo * Cout, (size_t)Ho * Wo * Cout, g.pitch, g.R, 1));
            bd->my[1] = bd->my[2] = bd->my[3] = bd->my[0];
        } else {   // dy is the HIGH-resolution gradient [cap, 2*Ho, 2*Wo, Cout]; variant (py,px) reads its parity view
            const int Wh = 2 * Wo, Hh = 2 * Ho;
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px)
                    RFV_TRY(make_map4(&bd->my[py * 2 + px], dy + ((size_t)py * Wh + px) * Cout, Cout, Wo, Ho, cap, (size_t)2 * Cout,
                                      (size_t)2 * Wh * Cout, (size_t)Hh * Wh * Cout, g.pitch, g.R, 1));
        }
        bd->smem = wgrad_smem_bytes(g);
        const int taps = kind == 1 ? 1 : 9;
        const double fl = 2.0 * taps * a->C * Cout * Ho * Wo * (kind == 3 ? 4 : 1);  // algorithmic (kind 3: at the high resolution)
        const int sms = num_sms;
        if (!dst_override) mark_grad(pi);
        push("wgrad_umma", "bwd:wgrad:" + label, fl, [this, bd, pi, sms, dst_override](const RunCtx& rc, cudaStream_t s) {
            WgradGeom g = bd->g;
            g.num_tiles = rc.B * g.tiles_per_img;
            const long long total = (long long)g.nvar * g.cchA * g.cchB * g.num_tiles;
            const int grid = (int)std::min<long long>(total, sms);
            wgrad_umma_kernel<<<grid, WG_THREADS, bd->smem, s>>>(bd->ma[0], bd->ma[1], bd->ma[2], bd->ma[3], bd->my[0], bd->my[1], bd->my[2],
                                                                 bd->my[3], dst_override ? dst_override : gslot(pi), g);
            return cudaGetLastError();
        });
        return 0;
    }

    // per-(image, channel) sums of dy into out_nc (row stride ld_nc; may be null) and per-channel sums into up to two
    // parameter-gradient slots (pi1 / pi2, -1: none)
    void bwd_colsum(const std::string& label, const bf16* dy, int C, int HW, float* out_nc, int ld_nc, int pi1, int pi2) {
        const int vpp = C / 8;
        const int threads = (256 / vpp) * vpp;
        const int ppb = std::min(HW, std::max(1, 32768 / C));
        mark_grad(pi1);
        mark_grad(pi2);
        push("colsum", "bwd:colsum:" + label, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
            dim3 grid((HW + ppb - 1) / ppb, rc.B);
            colsum_kernel<<<grid, threads, C * sizeof(float), s>>>(dy, out_nc, ld_nc, pi1 >= 0 ? gslot(pi1) : nullptr,
                                                                   pi2 >= 0 ? gslot(pi2) : nullptr, C, HW, ppb);
            return cudaGetLastError();
        });
    }

    // GroupNorm(+SiLU)(+dropout) backward of `site`: gradients of its sources (+)= f(dy) + addends.
    // ka / kb: consumer index this norm's block holds on srcs[0] / srcs[1]; out_override: write the (single-source)
    // result there instead of srcs[0]->grad (intermediate tensors that have no gradient buffer of their own).
    int bwd_gn(const std::string& label, const NormSite& st, const bf16* dy, const bf16* add_cat, const bf16* add_a, int ka, int kb,
               bf16* out_override = nullptr, float* out_colsum = nullptr, int ld_colsum = 0) {
        GnBwdArgs a{};
        a.dy = dy;
        a.xa = st.srcs[0]->p; a.stats_a = st.srcs[0]->stats; a.Ca = st.srcs[0]->C;
        if (st.srcs.size() > 1) { a.xb = st.srcs[1]->p; a.stats_b = st.srcs[1]->stats; a.Cb = st.srcs[1]->C; }
        a.gamma = pf(st.ig); a.beta = pf(st.ib);
        a.cs = st.cs;
        a.add_cat = add_cat; a.add_a = add_a;
        if (out_override) a.out_a = out_override;
        else {
            RFV_TRY(ensure_grad(st.srcs[0]));
            a.out_a = st.srcs[0]->grad;
            if (st.srcs.size() > 1) { RFV_TRY(ensure_grad(st.srcs[1])); a.out_b = st.srcs[1]->grad; }
        }
        a.HW = st.HW; a.slab_shift = slab_shift; a.silu = st.silu ? 1 : 0; a.eps = 1e-5f;
        a.out_colsum = out_colsum; a.ld_colsum = ld_colsum;
        const int C = st.C, vpp = C / 8, threads = (256 / vpp) * vpp;
        a.pix_per_block = std::min(st.HW, std::max(1, 65536 / C));
        const bool drop = st.drop;
        const int id = st.id, ig = st.ig, ib = st.ib;
        mark_grad(ig);
        mark_grad(ib);
        ActP sa = st.srcs[0], sb = st.srcs.size() > 1 ? st.srcs[1] : nullptr;
        const bool over = out_override != nullptr;
        auto fill = [=](GnBwdArgs& q, const RunCtx& rc) {
            q.dgamma = gslot(ig); q.dbeta = gslot(ib);
            q.acc_a = (!over && ka != sa->consumers - 1) ? 1 : 0;
            q.acc_b = (sb && kb != sb->consumers - 1) ? 1 : 0;
            q.drop_thresh = drop ? rc.drop_thresh : 0u;
            q.seed = rc.seed ^ ((uint32_t)id * 0x9E3779B9u);
            q.drop_scale = rc.drop_scale;
        };
        // Single-pass kernel: smallest number of pixel slices per image whose x + dy fit ~1/3 of an SM's shared memory.
        // Only for tensors of up to 128 channels: measured at 128 images (B200, one stream) 0.087 ms against 0.132 ms for the
        // two passes at 64 ch x 64x64, but 0.48 against 0.35 ms on the 192-channel concat (64-pixel slices: the per-CTA
        // prologue dominates) -- and its 3 x 74 KB of shared memory per SM displaces the weight-gradient kernel of the side
        // stream, so the whole step (two streams) goes 9.92 -> 9.81 ms with this limit and 9.92 -> 10.24 ms without it.
        int SL = 0;
        if (!(cfg.flags & RFV_FLAG_GN_BWD_TWO_PASS) && C <= 128) {
            for (int c = 1; c <= st.HW / GNF_STAGES; c *= 2) {
                if (st.HW % (c * GNF_STAGES) != 0) break;
                if (gn_bwd_fused_smem(C, st.HW / c, threads) <= 75264) { SL = c; break; }   // 3 CTAs per SM
            }
        }
        if (SL) {
            a.pix_per_block = st.HW / SL;
            a.arrive = reinterpret_cast<uint32_t*>(st.cs + (size_t)cap * C * 2);
            a.ticket = a.arrive + cap;
            const size_t smem = gn_bwd_fused_smem(C, a.pix_per_block, threads);
            push("gn_bwd", "bwd:gn_fused:" + label, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
                GnBwdArgs q = a;
                fill(q, rc);
                gn_bwd_fused_kernel<<<SL * rc.B, threads, smem, s>>>(q, SL);
                return cudaGetLastError();
            });
            set_last_bytes(6.0 * C * st.HW);   // x and dy in, dx out (addends extra)
            return 0;
        }
        push("gn_bwd", "bwd:gn_reduce:" + label, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
            GnBwdArgs q = a;
            fill(q, rc);
            dim3 grid((q.HW + q.pix_per_block - 1) / q.pix_per_block, rc.B);
            gn_bwd_kernel<false><<<grid, threads, 2 * C * sizeof(float), s>>>(q);
            return cudaGetLastError();
        });
        set_last_bytes(4.0 * C * st.HW);    // pass 1: x and dy in
        push("gn_bwd", "bwd:gn_apply:" + label, 0.0, [=](const RunCtx& rc, cudaStream_t s) {
            GnBwdArgs q = a;
            fill(q, rc);
            dim3 grid((q.HW + q.pix_per_block - 1) / q.pix_per_block, rc.B);
            gn_bwd_kernel<true><<<grid, threads, q.out_colsum ? C * sizeof(float) : 0, s>>>(q);
            return cudaGetLastError();
        });
        set_last_bytes(6.0 * C * st.HW);    // pass 2: x and dy in, dx out (addends extra)
        return 0;
    }

    int build();
    int finish_training_setup();
    int run_backward(const RunCtx& rc, cudaStream_t s, bool record_buckets = false);
    int run_forward(const RunCtx& rc, cudaStream_t s);
};

// ---------------------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------------------
int rfv_engine::build() {
    const int S = cfg.image_size, mc = cfg.model_channels, nlev = cfg.num_levels, nres = cfg.num_res_blocks;
    // opt in to > 48 KB of dynamic shared memory first: the cluster-occupancy queries below depend on it
    CU_CHECK(cudaFuncSetAttribute(conv_umma2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<128>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_wa_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU_CHECK(cudaFuncSetAttribute(conv_wa_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU_CHECK(cudaFuncSetAttribute(conv_wa_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU_CHECK(cudaFuncSetAttribute(conv_wa_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (train) CU_CHECK(cudaFuncSetAttribute(gn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 75264));
    td = 4 * mc;
    slab_shift = ilog2(mc / 8);
    std::vector<int> chans(nlev);
    for (int i = 0; i < nlev; ++i) chans[i] = mc * cfg.channel_mult[i];

    // concatenated time-projection matrix: one row block per ResidualBlock, in forward order
    sumC = 0;
    for (int lv = 0; lv < nlev; ++lv) sumC += nres * chans[lv];        // encoder
    sumC += 2 * chans[nlev - 1];                                       // middle
    for (int lv = 0; lv < nlev; ++lv) sumC += nres * chans[lv];        // decoder
    RFV_TRY(dalloc(&wcat, (size_t)sumC * td));
    RFV_TRY(dalloc(&bcat, (size_t)sumC));
    RFV_TRY(dalloc(&temb_act, (size_t)cap * td));
    RFV_TRY(dalloc(&tproj, (size_t)cap * sumC));
    RFV_TRY(dalloc(&t_steps, (size_t)cap));
    {   // every stats-carrying tensor holds cap x (C / slab) x 2 floats with C / slab = 8 * channel_mult
        int mmax = 1;
        for (int i = 0; i < nlev; ++i) mmax = std::max(mmax, cfg.channel_mult[i]);
        stats_floats = (size_t)cap * 16 * mmax * (4 * nlev * nres + 2 * nlev + 8);
    }
    RFV_TRY(dalloc(&stats_arena, stats_floats));
    if (train) {
        int cmax = 0;
        size_t need = 0;  // largest per-image backward temporary (elements)
        for (int lv = 0; lv < nlev; ++lv) {
            const size_t hw = (size_t)(S >> lv) * (S >> lv);
            const int cin_max = chans[lv] + (lv + 1 < nlev ? chans[lv + 1] : chans[lv]);  // decoder concat [h | skip]
            need = std::max(need, hw * cin_max);
            if (lv > 0) need = std::max(need, hw * 4 * chans[lv]);                        // upsampled tensor of level lv at level lv-1
            cmax = std::max(cmax, cin_max);
        }
        need = std::max(need, (size_t)(S >> (nlev - 1)) * (S >> (nlev - 1)) * 3 * chans[nlev - 1]);  // d(qkv)
        cmax = std::max(cmax, 3 * chans[nlev - 1]);
        scratch_elems = need;
        for (int i = 0; i < 3; ++i) RFV_TRY(dalloc(&scratch[i], (size_t)cap * scratch_elems));
        // one [cap][C][2] slice per norm site: 2 per block (C_in + C_out <= 2 * cmax), attention, output
        cs_floats = (size_t)cap * 2 * ((size_t)(2 * nlev * nres + 2) * 2 * cmax + 2 * cmax) + (size_t)(cap + 4) * (4 * nlev * nres + 8);
        RFV_TRY(dalloc(&cs_arena, cs_floats));
        RFV_TRY(dalloc(&d_tproj, (size_t)cap * sumC));
        RFV_TRY(dalloc(&temb_emb, (size_t)cap * mc));
        RFV_TRY(dalloc(&temb_z1, (size_t)cap * td));
        RFV_TRY(dalloc(&temb_h1, (size_t)cap * td));
        RFV_TRY(dalloc(&temb_z2, (size_t)cap * td));
        RFV_TRY(dalloc(&d_tz2, (size_t)cap * td));
        RFV_TRY(dalloc(&d_tz1, (size_t)cap * td));
        RFV_TRY(dalloc(&dv_buf, (size_t)cap * cfg.out_channels * S * S));
        const int nlow = (S >> (nlev - 1)) * (S >> (nlev - 1));
        RFV_TRY(dalloc(&lse, (size_t)cap * cfg.num_heads * nlow));
        RFV_TRY(dalloc(&delta, (size_t)cap * cfg.num_heads * nlow));
        RFV_TRY(dalloc(&zero_bias, (size_t)std::max(cmax, 1024)));
        CU_CHECK(cudaMemset(zero_bias, 0, (size_t)std::max(cmax, 1024) * sizeof(float)));
        RFV_TRY(dalloc(&norm2, 1));
        RFV_TRY(dalloc(&norm_partial, SUMSQ_BLOCKS));
    }

    // ---- time embedding (models/unet.py:157-162,231) ----
    int tw1, tb1, tw2, tb2;
    RFV_TRY(add_param("time_mlp.1.weight", (int64_t)td * mc, &tw1));
    RFV_TRY(add_param("time_mlp.1.bias", td, &tb1));
    RFV_TRY(add_param("time_mlp.3.weight", (int64_t)td * td, &tw2));
    RFV_TRY(add_param("time_mlp.3.bias", td, &tb2));
    {
        float *w1 = pf(tw1), *b1 = pf(tb1), *w2 = pf(tw2), *b2 = pf(tb2), *act = temb_act, *proj = tproj, *wc = wcat, *bc = bcat;
        const int td_ = td, mc_ = mc, sumC_ = sumC;
        push("temb", "temb:time_mlp", 2.0 * (mc * td + td * td), [=](const RunCtx& rc, cudaStream_t s) {
            const int rows = rc.t ? rc.B : 1;
            float *se = rc.train ? temb_emb : nullptr, *s1 = rc.train ? temb_z1 : nullptr, *s2 = rc.train ? temb_z2 : nullptr;
            return klaunch(rc.pdl, temb_kernel, dim3(rows, rows >= 64 ? 1 : 8), dim3(256), (mc_ + td_) * sizeof(float), s, rc.t, rc.t ? 0 : 1, rc.t_scalar,
                           w1, b1, w2, b2, act, mc_, td_, se, s1, rc.train ? temb_h1 : nullptr, s2);
        });
        push("temb", "temb:block_projections", 2.0 * td * sumC, [=](const RunCtx& rc, cudaStream_t s) {
            const int rows = rc.t ? rc.B : 1;
            dim3 grid((sumC_ + 63) / 64, (rows + 7) / 8);
            return klaunch(rc.pdl, temb_proj_kernel, grid, dim3(256), 8 * td_ * sizeof(float), s, act, wc, bc, proj, rows, td_, sumC_);
        });
        if (train) {
            // backward of the time MLP: runs last (first block recorded), after every ResidualBlock has added its
            // d(time projection) rows into d_tproj.  Needs per-row t (training always passes a t vector).
            begin_bwd();
            // (the per-block projections are registered later: finish_training_setup marks them as final in this block)
            mark_grad(tw1); mark_grad(tb1); mark_grad(tw2); mark_grad(tb2);
            temb_block = (int)bwd_blocks.size() - 1;
            push("temb_bwd", "bwd:temb", 0.0, [=](const RunCtx& rc, cudaStream_t s) {
                const int B = rc.B;
                // dW_block = d_tproj[:, off:off+C]^T . temb_act ; biases (+ conv1.bias) = column sums: all blocks in one launch
                LinSegs sg{};
                if ((int)time_projs.size() > LIN_MAX_SEGS) return cudaErrorInvalidValue;
                for (auto& tp : time_projs) sg.s[sg.n++] = LinSeg{tp.off, tp.Cout, gslot(tp.iw), gslot(tp.ib), gslot(tp.icb)};
                lin_wgrad_kernel<<<(sumC_ * td_ + 255) / 256, 256, 0, s>>>(d_tproj, sumC_, act, td_, sg, B, sumC_, td_);
                // dz2 = (d_tproj . Wcat) * silu'(z2)
                lin_dgrad_kernel<<<(B * td_ + 255) / 256, 256, 0, s>>>(d_tproj, sumC_, wc, d_tz2, temb_z2, B, sumC_, td_);
                LinSegs s2{};
                s2.n = 1; s2.s[0] = LinSeg{0, td_, gslot(tw2), gslot(tb2), nullptr};
                lin_wgrad_kernel<<<(td_ * td_ + 255) / 256, 256, 0, s>>>(d_tz2, td_, temb_h1, td_, s2, B, td_, td_);
                // dz1 = (dz2 . W2) * silu'(z1)
                lin_dgrad_kernel<<<(B * td_ + 255) / 256, 256, 0, s>>>(d_tz2, td_, w2, d_tz1, temb_z1, B, td_, td_);
                LinSegs s1{};
                s1.n = 1; s1.s[0] = LinSeg{0, td_, gslot(tw1), gslot(tb1), nullptr};
                lin_wgrad_kernel<<<(td_ * mc_ + 255) / 256, 256, 0, s>>>(d_tz1, td_, temb_emb, mc_, s1, B, td_, mc_);
                return cudaGetLastError();
            });
            end_bwd();
        }
    }

    // ---- input conv (models/unet.py:165,234) ----
    ActP h;
    {
        int iw, ib;
        RFV_TRY(add_param("input_conv.weight", (int64_t)mc * cfg.in_channels * 9, &iw));
        RFV_TRY(add_param("input_conv.bias", mc, &ib));
        RFV_TRY(new_act(&h, mc, S, S, true));
        const int Cin = cfg.in_channels, ss = slab_shift;
        const int K = Cin * 9;
        float* wt = nullptr;  // [K][mc] fp32: the layout the kernel stages into shared memory
        RFV_TRY(dalloc(&wt, (size_t)K * mc));
        params[iw].repack = [this, iw, wt, K, mc](cudaStream_t s) {
            transpose_input_weight_kernel<<<(K * mc + 255) / 256, 256, 0, s>>>(pf(iw), wt, mc, K);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "input weight transpose failed");
        };
        const float* b = pf(ib);
        bf16* o = h->p;
        float* st = h->stats;
        const size_t smem = ((size_t)K * mc + mc + (mc / 8) * 2) * sizeof(float);
        if (S % 2 != 0) return fail(RFV_ERR_INVALID, "image_size must be even");
        // tensor-core version where the channel count splits into 64- (or one 32-) channel chunks
        const int im_ntc = (mc % 64 == 0) ? 8 : (mc == 32 ? 4 : 0);
        const bool im = im_ntc != 0 && !(cfg.flags & RFV_FLAG_INPUT_CONV_FMA);
        const size_t im_smem = ((size_t)Cin * (IM_TH + 2) * IM_XP + (size_t)((((mc >> ss) * 2) + 3) & ~3)) * sizeof(float) +
                               (size_t)8 * 32 * (im_ntc * 16 + 16);
        push("input_conv", "conv:input_conv", 2.0 * K * mc * S * S, [=](const RunCtx& rc, cudaStream_t s) {
            if (im) {
                const int tiles_img = ((S + IM_TW - 1) / IM_TW) * ((S + IM_TH - 1) / IM_TH);
                dim3 grid((tiles_img + IM_TPB - 1) / IM_TPB, rc.B);
#define RFV_IM_LAUNCH(CI, NT) return klaunch(rc.pdl, input_conv_mma_kernel<CI, NT>, grid, dim3(256), im_smem, s, rc.x, rc.x1, rc.t, wt, b, o, st, S, S, mc, ss)
                if (im_ntc == 8) {
                    switch (Cin) {
                        case 1: RFV_IM_LAUNCH(1, 8); break;
                        case 2: RFV_IM_LAUNCH(2, 8); break;
                        case 3: RFV_IM_LAUNCH(3, 8); break;
                        default: RFV_IM_LAUNCH(4, 8); break;
                    }
                } else {
                    switch (Cin) {
                        case 1: RFV_IM_LAUNCH(1, 4); break;
                        case 2: RFV_IM_LAUNCH(2, 4); break;
                        case 3: RFV_IM_LAUNCH(3, 4); break;
                        default: RFV_IM_LAUNCH(4, 4); break;
                    }
                }
#undef RFV_IM_LAUNCH
                return cudaGetLastError();
            }
            dim3 grid((S * S + 511) / 512, rc.B);
            switch (Cin) {
                case 1: input_conv_kernel<1><<<grid, 256, smem, s>>>(rc.x, rc.x1, rc.t, wt, b, o, st, S, S, mc, ss); break;
                case 2: input_conv_kernel<2><<<grid, 256, smem, s>>>(rc.x, rc.x1, rc.t, wt, b, o, st, S, S, mc, ss); break;
                case 3: input_conv_kernel<3><<<grid, 256, smem, s>>>(rc.x, rc.x1, rc.t, wt, b, o, st, S, S, mc, ss); break;
                default: input_conv_kernel<4><<<grid, 256, smem, s>>>(rc.x, rc.x1, rc.t, wt, b, o, st, S, S, mc, ss); break;
            }
            return cudaGetLastError();
        });
        named_acts["input_conv"] = h;
        if (train) {
            RFV_TRY(ensure_grad(h));
            float* wscr = nullptr;   // padded weight gradient [mc][9][64]
            RFV_TRY(dalloc(&wscr, (size_t)mc * 9 * 64));
            begin_bwd();
            const bf16* dh = h->grad;
            ActP xpad;
            RFV_TRY(scratch_act(&xpad, 0, 64, S, S));
            {
                bf16* xp = xpad->p;
                push("elementwise_bwd", "bwd:pad:input_conv", 0.0, [=](const RunCtx& rc, cudaStream_t s) {
                    cudaError_t e = cudaMemsetAsync(wscr, 0, (size_t)mc * 9 * 64 * sizeof(float), s);
                    if (e != cudaSuccess) return e;
                    pad_to_nhwc64_kernel<<<2048, 256, 0, s>>>(rc.x, rc.x1, rc.t, xp, rc.B, Cin, S * S);
                    return cudaGetLastError();
                });
            }
            on_side(true);
            RFV_TRY(bwd_wgrad("input_conv", 0, xpad, dh, mc, S, S, iw, 9 * 64, 0, wscr));
            mark_grad(iw);
            push("elementwise_bwd", "bwd:extract:input_conv", 0.0, [=](const RunCtx&, cudaStream_t s) {
                extract_wgrad_kernel<<<(mc * Cin * 9 + 255) / 256, 256, 0, s>>>(wscr, gslot(iw), mc, Cin, 64);
                return cudaGetLastError();
            });
            bwd_colsum("input_conv.bias", dh, mc, S * S, nullptr, 0, ib, -1);
            end_bwd();
        }
    }

    int temb_cursor = 0;
    std::vector<ActP> skips;
    int bi = 0, res = S;
    for (int lv = 0; lv < nlev; ++lv) {
        for (int r = 0; r < nres; ++r) {
            ActP out;
            RFV_TRY(res_block("enc_blocks." + std::to_string(bi), {h}, chans[lv], &temb_cursor, &out));
            release(h);
            h = out;
            ++bi;
        }
        retain(h);
        skips.push_back(h);
        if (lv < nlev - 1) {
            ConvLayer* d;
            RFV_TRY(add_conv(&d, "downsamples." + std::to_string(lv), chans[lv], chans[lv], 3, 2, 0, "", 0, 0));
            ActP o;
            res /= 2;
            RFV_TRY(new_act(&o, chans[lv], res, res, true));
            const int kd = h->consumers++;
            RFV_TRY(conv_op(d, h, {}, nullptr, o, -1, true));
            if (train) {
                RFV_TRY(ensure_grad(o));
                RFV_TRY(ensure_grad(h));
                ConvLayer* gd;
                RFV_TRY(add_dgrad_layer(&gd, d, 2));
                begin_bwd();
                on_side(false);
                RFV_TRY(bwd_wgrad(d->name, 2, h, o->grad, chans[lv], res, res, d->iw, 9 * chans[lv], 0));
                bwd_colsum(d->name + ".bias", o->grad, chans[lv], res * res, nullptr, 0, d->ib, -1);
                on_main();
                RFV_TRY(conv_op(gd, grad_view(o), {}, nullptr, grad_view(h), -1, false, h, kd));
                end_bwd();
            }
            release(h);
            h = o;
            named_acts["downsamples." + std::to_string(lv)] = h;
        }
    }
    // ---- middle ----
    {
        ActP out;
        RFV_TRY(res_block("mid_block1", {h}, h->C, &temb_cursor, &out));
        release(h);
        h = out;
        // attention (models/unet.py:79-100)
        const int C = h->C, N = res * res, heads = cfg.num_heads, d = C / heads;
        if (C % heads != 0 || (d != 32 && d != 64)) return fail(RFV_ERR_INVALID, "attention head dim %d unsupported (32 or 64)", d);
        if (N % 64 != 0) return fail(RFV_ERR_INVALID, "attention needs H*W %% 64 == 0 at the lowest level (got %d)", N);
        ActP hn, qkv, ao, o2;
        NormSite na;
        const int kat = h->consumers++;
        RFV_TRY(new_act(&hn, C, res, res, false));
        RFV_TRY(gn_op("mid_attn.norm", {h}, hn, false, false, &na));
        ConvLayer *cq, *cp;
        RFV_TRY(add_conv(&cq, "mid_attn.qkv", C, 3 * C, 1, 1, 0, "", 0, 0));
        RFV_TRY(new_act(&qkv, 3 * C, res, res, false));
        RFV_TRY(conv_op(cq, hn, {}, nullptr, qkv, -1, false));
        release(hn);
        RFV_TRY(new_act(&ao, C, res, res, false));
        {
            const bf16* qp = qkv->p;
            bf16* op = ao->p;
            const float sl2 = (1.0f / std::sqrt((float)d)) * 1.4426950408889634f;
            // tcgen05 kernel for the default shape (256 tokens, head dim 64); the mma.sync kernel covers the rest
            const bool au = use_umma && d == 64 && N == 256 && !(cfg.flags & RFV_FLAG_NO_ATTN_UMMA);
            // longer sequences (1024 tokens at 128x128): the same core looped over 256-key blocks with the online softmax
            const bool ak = use_umma && d == 64 && N > 256 && N % 256 == 0 && !(cfg.flags & RFV_FLAG_NO_ATTN_UMMA);
            auto amap = std::make_shared<CUtensorMap>();
            if (au || ak) {
                RFV_TRY(make_map2(amap.get(), qkv->p, 3 * C, cap * N, 128));
                CU_CHECK(cudaFuncSetAttribute(attn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AU_SMEM));
                CU_CHECK(cudaFuncSetAttribute(attn_umma_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AK_SMEM));
            }
            push("attention", "attn:mid_attn", 4.0 * C * (double)N * N, [=](const RunCtx& rc, cudaStream_t s) {
                float* l = rc.train ? lse : nullptr;
                if (au) {
                    return klaunch(rc.pdl, attn_umma_kernel, dim3(N / 128, heads, rc.B), dim3(AU_THREADS), AU_SMEM, s, *amap, op, N, C, sl2, l);
                }
                if (ak) {
                    return klaunch(rc.pdl, attn_umma_kv_kernel, dim3(N / 128, heads, rc.B), dim3(AK_THREADS), AK_SMEM, s, *amap, op, N, C, sl2, l);
                }
                dim3 grid(N / 64, heads, rc.B);
                if (d == 64) return klaunch(rc.pdl, attn_kernel<64>, grid, dim3(128), 0, s, qp, op, N, C, sl2, l);
                return klaunch(rc.pdl, attn_kernel<32>, grid, dim3(128), 0, s, qp, op, N, C, sl2, l);
            });
        }
        release(qkv);
        RFV_TRY(add_conv(&cp, "mid_attn.proj", C, C, 1, 1, 0, "", 0, 0));
        RFV_TRY(new_act(&o2, C, res, res, true));
        RFV_TRY(conv_op(cp, ao, {}, h, o2, -1, true));
        if (train) {
            // ---- backward of the attention block (models/unet.py:79-100 reversed) ----
            RFV_TRY(ensure_grad(o2));
            ConvLayer *gp, *gq;
            RFV_TRY(add_dgrad_layer(&gp, cp, 0));
            RFV_TRY(add_dgrad_layer(&gq, cq, 0));
            begin_bwd();
            ActP T0, T1, T2;
            on_side(false);
            RFV_TRY(bwd_wgrad("mid_attn.proj", 1, ao, o2->grad, C, res, res, cp->iw, C, 0));
            bwd_colsum("mid_attn.proj.bias", o2->grad, C, N, nullptr, 0, cp->ib, -1);
            on_main();
            RFV_TRY(scratch_act(&T0, 0, C, res, res));
            RFV_TRY(conv_op(gp, grad_view(o2), {}, nullptr, T0, -1, false));   // d(attention output)
            RFV_TRY(scratch_act(&T1, 1, 3 * C, res, res));
            {
                const bf16 *qp = qkv->p, *op = ao->p, *dop = T0->p;
                bf16* dq = T1->p;
                const float sc = 1.0f / std::sqrt((float)d), sl2 = sc * 1.4426950408889634f;
                push("attention_bwd", "bwd:attn:mid_attn", 10.0 * C * (double)N * N, [=](const RunCtx& rc, cudaStream_t s) {
                    dim3 grid(N / 64, heads, rc.B);
                    if (d == 64) {
                        attn_bwd_dq_kernel<64><<<grid, 128, 0, s>>>(qp, op, dop, lse, delta, dq, N, C, sc, sl2);
                        attn_bwd_dkv_kernel<64><<<grid, 128, 0, s>>>(qp, dop, lse, delta, dq, N, C, sc, sl2);
                    } else {
                        attn_bwd_dq_kernel<32><<<grid, 128, 0, s>>>(qp, op, dop, lse, delta, dq, N, C, sc, sl2);
                        attn_bwd_dkv_kernel<32><<<grid, 128, 0, s>>>(qp, dop, lse, delta, dq, N, C, sc, sl2);
                    }
                    return cudaGetLastError();
                });
            }
            on_side(true);    // reads d(qkv), which main has just produced
            RFV_TRY(bwd_wgrad("mid_attn.qkv", 1, hn, T1->p, 3 * C, res, res, cq->iw, C, 0));
            bwd_colsum("mid_attn.qkv.bias", T1->p, 3 * C, N, nullptr, 0, cq->ib, -1);
            on_main();
            RFV_TRY(scratch_act(&T2, 2, C, res, res));
            RFV_TRY(conv_op(gq, T1, {}, nullptr, T2, -1, false));              // d(normalised input)
            RFV_TRY(bwd_gn("mid_attn.norm", na, T2->p, nullptr, o2->grad, kat, 0));  // + identity path x + h
            end_bwd();
        }
        release(ao);
        release(h);
        h = o2;
        named_acts["mid_attn"] = h;
        RFV_TRY(res_block("mid_block2", {h}, h->C, &temb_cursor, &out));
        release(h);
        h = out;
    }
    // ---- decoder ----
    bi = 0;
    for (int li = 0; li < nlev; ++li) {
        const int lv = nlev - 1 - li;
        ActP sk = skips.back();
        skips.pop_back();
        ActP out;
        RFV_TRY(res_block("dec_blocks." + std::to_string(bi), {h, sk}, chans[lv], &temb_cursor, &out));
        release(h);
        release(sk);  // the reference taken when the skip was pushed
        h = out;
        ++bi;
        for (int r = 0; r < nres - 1; ++r) {
            RFV_TRY(res_block("dec_blocks." + std::to_string(bi), {h}, chans[lv], &temb_cursor, &out));
            release(h);
            h = out;
            ++bi;
        }
        if (lv > 0) {
            ConvLayer* u;
            RFV_TRY(add_conv(&u, "upsamples." + std::to_string(li) + ".1", chans[lv], chans[lv], 3, 1, 1, "", 0, 0));
            ActP o;
            res *= 2;
            RFV_TRY(new_act(&o, chans[lv], res, res, true));
            const int ku = h->consumers++;
            RFV_TRY(conv_op(u, h, {}, nullptr, o, -1, true));
            if (train) {
                // backward of nearest-x2 + conv3x3 (models/unet.py:215-218), both in the sub-pixel formulation of the forward:
                // nothing is upsampled or pooled, 2.25x fewer MACs than the literal high-resolution 3x3
                RFV_TRY(ensure_grad(o));
                RFV_TRY(ensure_grad(h));
                ConvLayer* gu;
                RFV_TRY(add_dgrad_layer(&gu, u, 3));
                begin_bwd();
                const int Cc = chans[lv], lo = res / 2;
                // weight gradient in the sub-pixel formulation (four phases x 2x2 taps over the LOW-resolution input; each
                // pre-summed tap's gradient is added to the 3x3 taps it stands for): 2.25x fewer MACs, nothing materialised
                on_side(false);
                RFV_TRY(bwd_wgrad(u->name, 3, h, o->grad, Cc, lo, lo, u->iw, 9 * Cc, 0));
                bwd_colsum(u->name + ".bias", o->grad, Cc, res * res, nullptr, 0, u->ib, -1);
                on_main();
                RFV_TRY(conv_op(gu, grad_view(o), {}, nullptr, grad_view(h), -1, false, h, ku));
                end_bwd();
            }
            release(h);
            h = o;
            named_acts["upsamples." + std::to_string(li)] = h;
        }
    }
    // ---- output (models/unet.py:223-227,275) fused with the Euler update (models/base_flow.py:170) ----
    {
        ActP a;
        NormSite no;
        const int ko = h->consumers++;
        ActP hin = h;
        RFV_TRY(new_act(&a, h->C, S, S, false));
        RFV_TRY(gn_op("output_conv.0", {h}, a, true, false, &no));
        release(h);
        int iw, ib;
        const int C = a->C, Co = cfg.out_channels;
        RFV_TRY(add_param("output_conv.2.weight", (int64_t)Co * C * 9, &iw));
        RFV_TRY(add_param("output_conv.2.bias", Co, &ib));
        if (Co > 4) return fail(RFV_ERR_INVALID, "out_channels > 4 unsupported");
        bf16* wpk = nullptr;  // [9 taps][8 (C_out padded)][C] bf16: B operand of the mma.sync formulation
        RFV_TRY(dalloc(&wpk, (size_t)9 * 8 * C));
        params[iw].repack = [this, iw, wpk, Co, C](cudaStream_t s) {
            pack_output_weight_kernel<<<(9 * 8 * C + 255) / 256, 256, 0, s>>>(pf(iw), wpk, Co, C);
            return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "output weight pack failed");
        };
        const float* b = pf(ib);
        const bf16* ap = a->p;
        const int pitch = C * 2 + 16;
        const size_t smem = (size_t)2 * (OC_TH + 2) * (OC_TW + 2) * pitch + (size_t)9 * 8 * pitch;
        if (smem > 227 * 1024) return fail(RFV_ERR_INVALID, "output conv: %d channels do not fit the shared-memory tile", C);
        CU_CHECK(cudaFuncSetAttribute(output_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int sms = num_sms;
        // multiply-once formulation where it applies (C_out <= 3, 32 / 64 / 128 channels)
        const bool oz = Co <= 3 && (C == 32 || C == 64 || C == 128) && !(cfg.flags & RFV_FLAG_OUTPUT_CONV_TAPS);
        const size_t zsmem = output_conv_z_smem(C);
        if (oz) {
            CU_CHECK(cudaFuncSetAttribute(output_conv_z_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)output_conv_z_smem(32)));
            CU_CHECK(cudaFuncSetAttribute(output_conv_z_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)output_conv_z_smem(64)));
            CU_CHECK(cudaFuncSetAttribute(output_conv_z_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)output_conv_z_smem(128)));
        }
        push("output_conv", "conv:output_conv.2", 2.0 * 9 * C * Co * S * S, [=](const RunCtx& rc, cudaStream_t s) {
            const int ntiles = ((S + OC_TW - 1) / OC_TW) * ((S + OC_TH - 1) / OC_TH) * rc.B;
            const int grid = std::min(ntiles, 2 * sms);
            if (oz) {
                auto kz = C == 32 ? output_conv_z_kernel<2> : (C == 64 ? output_conv_z_kernel<4> : output_conv_z_kernel<8>);
                return klaunch(rc.pdl, kz, dim3(grid), dim3(256), zsmem, s, ap, wpk, b, rc.out, rc.traj, rc.tgt_x0, rc.tgt_x1, rc.mse, S, S, Co, rc.B, rc.mode, rc.dt);
            }
            return klaunch(rc.pdl, output_conv_kernel, dim3(grid), dim3(256), smem, s, ap, wpk, b, rc.out, rc.traj, rc.tgt_x0, rc.tgt_x1, rc.mse, C, S, S, Co, rc.B,
                           rc.mode, rc.dt);
        });
        if (train) {
            // backward of the output conv: dv (fp32 NCHW, written by the forward in mode 3) -> weight / bias gradient,
            // data gradient through the thin-conv kernel on flipped weights, then the GroupNorm backward
            float* wdg = nullptr;  // [Co*9][C] fp32
            RFV_TRY(dalloc(&wdg, (size_t)Co * 9 * C));
            {
                auto prev = params[iw].repack;
                params[iw].repack = [this, prev, iw, wdg, Co, C](cudaStream_t s) {
                    const int rc = prev ? prev(s) : 0;
                    if (rc) return rc;
                    pack_output_dgrad_weight_kernel<<<(Co * 9 * C + 255) / 256, 256, 0, s>>>(pf(iw), wdg, Co, C);
                    return cudaGetLastError() == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "output dgrad weight pack failed");
                };
            }
            begin_bwd();
            ActP T0;
            RFV_TRY(scratch_act(&T0, 0, C, S, S));
            const size_t ic_smem = ((size_t)Co * 9 * C + C + (C / 8) * 2) * sizeof(float);
            bf16* t0 = T0->p;
            const int ss = slab_shift;
            float* wscr = nullptr;   // padded weight gradient [64][9][C]
            RFV_TRY(dalloc(&wscr, (size_t)64 * 9 * C));
            ActP dvpad;
            RFV_TRY(scratch_act(&dvpad, 1, 64, S, S));
            {
                bf16* dp = dvpad->p;
                mark_grad(ib);
                push("elementwise_bwd", "bwd:pad:output_conv.2", 0.0, [=](const RunCtx& rc, cudaStream_t s) {
                    cudaError_t e = cudaMemsetAsync(wscr, 0, (size_t)64 * 9 * C * sizeof(float), s);
                    if (e != cudaSuccess) return e;
                    pad_to_nhwc64_kernel<<<2048, 256, 0, s>>>(dv_buf, nullptr, nullptr, dp, rc.B, Co, S * S);
                    nchw_channel_sum_kernel<<<dim3(64, Co), 256, 0, s>>>(dv_buf, gslot(ib), rc.B, Co, S * S);
                    return cudaGetLastError();
                });
            }
            on_side(true);
            RFV_TRY(bwd_wgrad("output_conv.2", 0, a, dvpad->p, 64, S, S, iw, 9 * C, 0, wscr));
            mark_grad(iw);
            push("elementwise_bwd", "bwd:extract:output_conv.2", 0.0, [=](const RunCtx&, cudaStream_t s) {
                extract_wgrad_kernel<<<(Co * C * 9 + 255) / 256, 256, 0, s>>>(wscr, gslot(iw), Co, C, C);
                return cudaGetLastError();
            });
            on_main();
            const int dg_ntc = (C % 64 == 0) ? 8 : (C == 32 ? 4 : 0);
            const bool dg_mma = dg_ntc != 0 && !(cfg.flags & RFV_FLAG_INPUT_CONV_FMA);
            const size_t dg_smem = ((size_t)Co * (IM_TH + 2) * IM_XP + (size_t)((((C >> ss) * 2) + 3) & ~3)) * sizeof(float) +
                                   (size_t)8 * 32 * (dg_ntc * 16 + 16);
            push("input_conv", "bwd:dgrad:output_conv.2", 2.0 * 9 * C * Co * S * S, [=](const RunCtx& rc, cudaStream_t s) {
                if (dg_mma) {   // same thin-K GEMM as the input conv: dv (fp32 NCHW, C_out channels) x flipped weights -> C channels
                    const int tiles_img = ((S + IM_TW - 1) / IM_TW) * ((S + IM_TH - 1) / IM_TH);
                    dim3 grid((tiles_img + IM_TPB - 1) / IM_TPB, rc.B);
#define RFV_DG_LAUNCH(CI, NT) input_conv_mma_kernel<CI, NT><<<grid, 256, dg_smem, s>>>(dv_buf, nullptr, nullptr, wdg, zero_bias, t0, nullptr, S, S, C, ss)
                    if (dg_ntc == 8) {
                        switch (Co) {
                            case 1: RFV_DG_LAUNCH(1, 8); break;
                            case 2: RFV_DG_LAUNCH(2, 8); break;
                            case 3: RFV_DG_LAUNCH(3, 8); break;
                            default: RFV_DG_LAUNCH(4, 8); break;
                        }
                    } else {
                        switch (Co) {
                            case 1: RFV_DG_LAUNCH(1, 4); break;
                            case 2: RFV_DG_LAUNCH(2, 4); break;
                            case 3: RFV_DG_LAUNCH(3, 4); break;
                            default: RFV_DG_LAUNCH(4, 4); break;
                        }
                    }
#undef RFV_DG_LAUNCH
                    return cudaGetLastError();
                }
                dim3 grid((S * S + 511) / 512, rc.B);
                switch (Co) {
                    case 1: input_conv_kernel<1><<<grid, 256, ic_smem, s>>>(dv_buf, nullptr, nullptr, wdg, zero_bias, t0, nullptr, S, S, C, ss); break;
                    case 2: input_conv_kernel<2><<<grid, 256, ic_smem, s>>>(dv_buf, nullptr, nullptr, wdg, zero_bias, t0, nullptr, S, S, C, ss); break;
                    case 3: input_conv_kernel<3><<<grid, 256, ic_smem, s>>>(dv_buf, nullptr, nullptr, wdg, zero_bias, t0, nullptr, S, S, C, ss); break;
                    default: input_conv_kernel<4><<<grid, 256, ic_smem, s>>>(dv_buf, nullptr, nullptr, wdg, zero_bias, t0, nullptr, S, S, C, ss); break;
                }
                return cudaGetLastError();
            });
            RFV_TRY(bwd_gn("output_conv.0", no, t0, nullptr, nullptr, ko, 0));
            end_bwd();
            (void)hin;
        }
        release(a);
    }
    if (temb_cursor != sumC) return fail(RFV_ERR_STATE, "internal: time-projection layout mismatch (%d vs %d)", temb_cursor, sumC);
    CU_CHECK(cudaFuncSetAttribute(conv_umma2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<128>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128>::SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(conv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64>::SMEM_BYTES));
    return 0;
}

int rfv_engine::run_forward(const RunCtx& rc_in, cudaStream_t s) {
    RunCtx rc = rc_in;
    // programmatic dependent launch for every kernel but the first of the pass (its predecessor is the memset node below)
    // Small micro-batches only: measured on one box, batch 64: 1.402 vs 1.428 ms per velocity evaluation (kernels of 10-30 us, the
    // launch gap and the prologue are a visible share); micro-batch 512: 8.45 vs 8.35 ms -- the dependents' CTAs hold the SMs they
    // land on while they wait, which costs more there than the prologues they hide.  RFV_PDL_MAX overrides the bound.
    static const int pdl_max = getenv("RFV_PDL_MAX") ? atoi(getenv("RFV_PDL_MAX")) : 128;
    const bool pdl_ok = use_pdl && !profiling && !rc.train && rc.B <= pdl_max;
    rc.pdl = false;
    for (auto& p : params)
        if (!p.loaded) return fail(RFV_ERR_STATE, "parameter %s was never uploaded (rfv_set_tensor)", p.name.c_str());
    if (rc.B < 1 || rc.B > cap) return fail(RFV_ERR_STATE, "micro-batch %d outside [1,%d]", rc.B, cap);
    have_fwd = false;   // every forward overwrites the activation arena (rfv_train_forward re-arms the flag afterwards)
    if (!rc.temb_only) CU_CHECK(cudaMemsetAsync(stats_arena, 0, stats_used * sizeof(float), s));
    while (profiling && prof_events.size() < ops.size()) {
        cudaEvent_t a, b;
        CU_CHECK(cudaEventCreate(&a));
        CU_CHECK(cudaEventCreate(&b));
        prof_events.push_back({a, b});
    }
    // Euler loops with a precomputed projection table skip the time-MLP ops; the table fill runs only them
    // RFV_ONLY_KIND=<kernel class> (diagnosis only, tools/power_by_kind.py): launch just that class, on whatever the buffers
    // hold -- the results are meaningless, the point is the board power / clock that class draws when run back to back
    static const char* only_kind = getenv("RFV_ONLY_KIND");
    auto skipped = [&](const Op& op) {
        if (only_kind && *only_kind && op.kind != only_kind) return true;
        const bool is_temb = op.kind == "temb";
        return rc.temb_only ? !is_temb : (is_temb && rc.temb_row >= 0);
    };
    for (size_t i = 0; i < ops.size(); ++i) {
        Op& op = ops[i];
        if (skipped(op)) continue;
        if (profiling) CU_CHECK(cudaEventRecord(prof_events[i].first, s));
        cudaError_t e = op.run(rc, s);
        if (e != cudaSuccess) return fail(RFV_ERR_CUDA, "launch of %s failed: %s", op.label.c_str(), cudaGetErrorString(e));
        ++launches;
        rc.pdl = pdl_ok;
        if (profiling) CU_CHECK(cudaEventRecord(prof_events[i].second, s));
    }
    if (profiling) {
        CU_CHECK(cudaStreamSynchronize(s));
        for (size_t i = 0; i < ops.size(); ++i) {
            if (skipped(ops[i])) continue;
            float ms = 0.f;
            CU_CHECK(cudaEventElapsedTime(&ms, prof_events[i].first, prof_events[i].second));
            auto& slot = prof[ops[i].kind + " " + ops[i].label];
            slot.first += ms;
            slot.second += 1;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// training: flat buffers, backward driver
// ---------------------------------------------------------------------------------------------------------
int rfv_engine::finish_training_setup() {
    // the time-MLP backward (recorded first, run last) also writes every block's projection weights / biases and conv1 biases
    for (auto& tp : time_projs)
        for (int pi : {tp.iw, tp.ib, tp.icb}) params[pi].final_block = std::min(params[pi].final_block, temb_block);
    for (auto& pr : params)
        if (pr.final_block == (1 << 30)) return fail(RFV_ERR_STATE, "internal: no backward op writes the gradient of %s", pr.name.c_str());
    {   // slots in completion order (stable within a block), then ~4 buckets of similar size cut at block boundaries
        std::vector<int> order(params.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return params[a].final_block > params[b].final_block; });
        int64_t off = 0;
        const int64_t target = gtotal / 4;
        GradBucket cur;
        for (size_t k = 0; k < order.size(); ++k) {
            Param& pr = params[order[k]];
            pr.goff = off;
            off += pr.numel;
            cur.numel += pr.numel;
            cur.ready_block = std::min<int>(pr.final_block, (int)bwd_blocks.size() - 1);
            const bool last = k + 1 == order.size();
            if (last || (cur.numel >= target && params[order[k + 1]].final_block != pr.final_block)) {
                buckets.push_back(cur);
                cur = GradBucket{};
                cur.off = off;
            }
        }
        for (auto& b : buckets)
            for (auto& e : b.ev) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    RFV_TRY(dalloc(&gflat, (size_t)gtotal));
    RFV_TRY(dalloc(&mflat, (size_t)gtotal));
    RFV_TRY(dalloc(&vflat, (size_t)gtotal));
    CU_CHECK(cudaMemset(gflat, 0, (size_t)gtotal * sizeof(float)));
    CU_CHECK(cudaMemset(mflat, 0, (size_t)gtotal * sizeof(float)));
    CU_CHECK(cudaMemset(vflat, 0, (size_t)gtotal * sizeof(float)));
    RFV_TRY(dalloc(&d_segs, params.size()));
    std::vector<int2> blocks;
    for (size_t i = 0; i < params.size(); ++i)
        for (int64_t c = 0; c * ADAM_CHUNK < params[i].numel; ++c) blocks.push_back(make_int2((int)i, (int)c));
    n_adam_blocks = (int)blocks.size();
    RFV_TRY(dalloc(&d_adam_blocks, blocks.size()));
    CU_CHECK(cudaMemcpy(d_adam_blocks, blocks.data(), blocks.size() * sizeof(int2), cudaMemcpyHostToDevice));
    CU_CHECK(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    {
        int least = 0, greatest = 0;
        CU_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU_CHECK(cudaStreamCreateWithPriority(&s_bwd, cudaStreamNonBlocking, greatest));
        CU_CHECK(cudaStreamCreateWithPriority(&s_side, cudaStreamNonBlocking, least));
        for (auto& e : ev_bwd) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    adam_dirty = true;
    return 0;
}

static AdamSeg make_seg(const Param& p) {
    AdamSeg s{};
    s.goff = p.goff; s.numel = p.numel; s.master = p.f32; s.bound = p.bound;
    s.O = p.O; s.I = p.I; s.KK = p.KK; s.ld = p.KK * p.I; s.koff = 0;
    return s;
}

int rfv_engine::run_backward(const RunCtx& rc, cudaStream_t s, bool record_buckets) {
    // hand over from the caller's stream to the engine's two backward streams (main: high priority, side: low)
    CU_CHECK(cudaEventRecord(ev_bwd[0], s));
    CU_CHECK(cudaStreamWaitEvent(s_bwd, ev_bwd[0], 0));
    CU_CHECK(cudaStreamWaitEvent(s_side, ev_bwd[0], 0));
    CU_CHECK(cudaMemsetAsync(cs_arena, 0, cs_used * sizeof(float), s_bwd));
    CU_CHECK(cudaMemsetAsync(d_tproj, 0, (size_t)cap * sumC * sizeof(float), s_bwd));
    for (size_t bi = bwd_blocks.size(); bi-- > 0;) {
        for (auto& op : bwd_blocks[bi]) {
            cudaStream_t st = (op.lane && two_streams) ? s_side : s_bwd;
            if (two_streams && (op.sync & 1)) {
                CU_CHECK(cudaEventRecord(ev_bwd[1], s_bwd));
                CU_CHECK(cudaStreamWaitEvent(s_side, ev_bwd[1], 0));
            }
            if (two_streams && (op.sync & 2)) {
                CU_CHECK(cudaEventRecord(ev_bwd[2], s_side));
                CU_CHECK(cudaStreamWaitEvent(s_bwd, ev_bwd[2], 0));
            }
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (profiling) {
                CU_CHECK(cudaEventCreate(&e0));
                CU_CHECK(cudaEventCreate(&e1));
                CU_CHECK(cudaEventRecord(e0, st));
            }
            cudaError_t e = op.run(rc, st);
            if (e != cudaSuccess) return fail(RFV_ERR_CUDA, "launch of %s failed: %s", op.label.c_str(), cudaGetErrorString(e));
            ++launches;
            if (profiling) {
                CU_CHECK(cudaEventRecord(e1, st));
                bwd_prof.push_back({&op, e0, e1});
            }
        }
        if (record_buckets)
            for (auto& b : buckets)
                if (b.ready_block == (int)bi) {   // everything that writes into this range is enqueued: one event per stream
                    CU_CHECK(cudaEventRecord(b.ev[0], s_bwd));
                    CU_CHECK(cudaEventRecord(b.ev[1], two_streams ? s_side : s_bwd));
                }
    }
    buckets_valid = record_buckets;
    CU_CHECK(cudaEventRecord(ev_bwd[2], s_side));
    CU_CHECK(cudaStreamWaitEvent(s_bwd, ev_bwd[2], 0));
    CU_CHECK(cudaEventRecord(ev_bwd[3], s_bwd));
    CU_CHECK(cudaStreamWaitEvent(s, ev_bwd[3], 0));
    if (profiling) {
        CU_CHECK(cudaStreamSynchronize(s));
        for (auto& r : bwd_prof) {
            float ms = 0.f;
            CU_CHECK(cudaEventElapsedTime(&ms, r.e0, r.e1));
            auto& slot = prof[r.op->kind + " " + r.op->label];
            slot.first += ms;
            slot.second += 1;
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
        bwd_prof.clear();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
RFV_EXPORT int rfv_abi_version(void) { return RFV_ABI_VERSION; }
RFV_EXPORT const char* rfv_last_error(void) { return g_err; }

RFV_EXPORT int rfv_create(const rfv_config* cfg, rfv_handle* out) {
    if (!cfg || !out) return fail(RFV_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_levels < 1 || cfg->num_levels > RFV_MAX_LEVELS) return fail(RFV_ERR_INVALID, "num_levels out of range");
    if (cfg->model_channels < 64 || cfg->model_channels % 64 != 0 || !is_pow2(cfg->model_channels / 8))
        return fail(RFV_ERR_INVALID, "model_channels must be 64, 128, 256, ... (got %d)", cfg->model_channels);
    if (cfg->in_channels < 1 || cfg->in_channels > 4 || cfg->out_channels < 1 || cfg->out_channels > 4)
        return fail(RFV_ERR_INVALID, "in/out channels must be in [1,4]");
    if (cfg->num_res_blocks < 1 || cfg->micro_batch < 1) return fail(RFV_ERR_INVALID, "num_res_blocks / micro_batch must be >= 1");
    const int down = 1 << (cfg->num_levels - 1);
    if (cfg->image_size % down != 0 || cfg->image_size / down < 8)
        return fail(RFV_ERR_INVALID, "image_size %d too small for %d levels (lowest level must be >= 8x8)", cfg->image_size, cfg->num_levels);
    CU_CHECK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(RFV_ERR_INVALID, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    auto e = std::make_unique<rfv_engine>();
    e->cfg = *cfg;
    e->num_sms = prop.multiProcessorCount;
    e->cap = (cfg->micro_batch + 1) & ~1;
    e->use_umma = !(cfg->flags & RFV_FLAG_NO_UMMA);
    e->keep_acts = (cfg->flags & RFV_FLAG_KEEP_ACTS) != 0;
    e->use_wa = !(cfg->flags & RFV_FLAG_NO_WA);
    e->use_lanes = !(cfg->flags & RFV_FLAG_ONE_LANE) && !(cfg->flags & RFV_FLAG_TRAIN);
    e->use_graphs = !(cfg->flags & RFV_FLAG_NO_GRAPH);
    e->use_pdl = !(cfg->flags & RFV_FLAG_NO_PDL) && !(cfg->flags & RFV_FLAG_TRAIN);
    e->two_streams = !(cfg->flags & RFV_FLAG_ONE_STREAM);
    e->fuse_mode = (cfg->flags & RFV_FLAG_FUSE_GN) ? 2 : ((cfg->flags & RFV_FLAG_NO_FUSE_GN) ? 0 : 1);
    e->train = (cfg->flags & RFV_FLAG_TRAIN) != 0;
    if (e->train) e->keep_acts = true;  // the backward pass reads every forward activation
    {
        const int c = (cfg->flags >> 8) & 7;
        if (c == 1 || c == 2 || c == 4) e->cluster = c;
        else if (c != 0) return fail(RFV_ERR_INVALID, "cluster size override must be 1, 2 or 4");
    }
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_weights, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    for (auto& ev : e->ev_join) CU_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(RFV_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
        e->encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    RFV_TRY(e->build());
    if (e->train) RFV_TRY(e->finish_training_setup());
    const size_t xin = (size_t)e->cap * cfg->in_channels * cfg->image_size * cfg->image_size;
    RFV_TRY(e->dalloc(&e->scratch_x, xin));
    RFV_TRY(e->dalloc(&e->xbuf[0], xin));
    RFV_TRY(e->dalloc(&e->xbuf[1], xin));
    CU_CHECK(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
    CU_CHECK(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
    CU_CHECK(cudaStreamCreateWithFlags(&e->s_cmp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU_CHECK(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
        CU_CHECK(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
        CU_CHECK(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
    }
    *out = e.release();
    return 0;
}

RFV_EXPORT int rfv_destroy(rfv_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    delete h;
    return 0;
}

RFV_EXPORT int rfv_num_tensors(rfv_handle h) { return h ? (int)h->params.size() : 0; }

RFV_EXPORT int rfv_tensor_info(rfv_handle h, int index, char* name_buf, int name_buf_len, int64_t* numel) {
    if (!h || index < 0 || index >= (int)h->params.size()) return fail(RFV_ERR_INVALID, "tensor index out of range");
    if (name_buf && name_buf_len > 0) snprintf(name_buf, name_buf_len, "%s", h->params[index].name.c_str());
    if (numel) *numel = h->params[index].numel;
    return 0;
}

RFV_EXPORT int rfv_set_tensor(rfv_handle h, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    if (!h || !name || !dev_ptr) return fail(RFV_ERR_INVALID, "null argument");
    auto it = h->param_index.find(name);
    if (it == h->param_index.end()) return fail(RFV_ERR_INVALID, "unknown tensor '%s'", name);
    Param& p = h->params[it->second];
    if (p.numel != numel) return fail(RFV_ERR_INVALID, "tensor '%s': expected %lld elements, got %lld", name, (long long)p.numel, (long long)numel);
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));   // kernels of an earlier call on another stream may still be reading these weights
    cudaError_t ce = cudaMemcpyAsync(p.f32, dev_ptr, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, s);
    int rc = ce == cudaSuccess ? 0 : fail(RFV_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(ce));
    if (rc == 0 && p.repack) rc = p.repack(s);
    if (rc == 0) {
        p.loaded = true;
        ce = cudaEventRecord(h->ev_weights, s);
        if (ce != cudaSuccess) rc = fail(RFV_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(ce));
    }
    const int rl = h->leave(s);   // also on failure: whatever was enqueued must order later calls
    if (rc == 0 && rl == 0 && h->lane) return rfv_set_tensor(h->lane, name, dev_ptr, numel, stream);
    return rc ? rc : rl;
}

RFV_EXPORT int rfv_get_tensor(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, void* stream) {
    if (!h || !name || !dev_ptr) return fail(RFV_ERR_INVALID, "null argument");
    auto it = h->param_index.find(name);
    if (it == h->param_index.end()) return fail(RFV_ERR_INVALID, "unknown tensor '%s'", name);
    Param& p = h->params[it->second];
    if (p.numel != numel) return fail(RFV_ERR_INVALID, "tensor '%s': expected %lld elements", name, (long long)p.numel);
    if (!p.loaded) return fail(RFV_ERR_STATE, "tensor '%s' not loaded", name);
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));   // an optimizer step enqueued on another stream may still be writing this tensor
    int rc = 0;
    if (p.readback) rc = p.readback(dev_ptr, s);
    else if (cudaMemcpyAsync(dev_ptr, p.f32, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        rc = fail(RFV_ERR_CUDA, "cudaMemcpyAsync failed");
    const int rl = h->leave(s);
    return rc ? rc : rl;
}

RFV_EXPORT int rfv_get_master(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, void* stream) {
    if (!h || !name || !dev_ptr) return fail(RFV_ERR_INVALID, "null argument");
    auto it = h->param_index.find(name);
    if (it == h->param_index.end()) return fail(RFV_ERR_INVALID, "unknown tensor '%s'", name);
    Param& p = h->params[it->second];
    if (p.numel != numel) return fail(RFV_ERR_INVALID, "tensor '%s': expected %lld elements", name, (long long)p.numel);
    if (!p.loaded) return fail(RFV_ERR_STATE, "tensor '%s' not loaded", name);
    RFV_TRY(h->enter((cudaStream_t)stream));
    int rc = 0;
    if (cudaMemcpyAsync(dev_ptr, p.f32, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess)
        rc = fail(RFV_ERR_CUDA, "cudaMemcpyAsync failed");
    const int rl = h->leave((cudaStream_t)stream);
    return rc ? rc : rl;
}

// dropout seed of micro-batch `idx` of a call: both halves of the caller's seed and the chunk index go through an avalanche
// mix, so consecutive seeds (the trainer passes its step counter) give unrelated mask streams, not XOR-permutations of one
static uint32_t mix_seed(uint64_t seed, uint32_t idx) {
    auto mix = [](uint32_t v) {
        v ^= v >> 16; v *= 0x85ebca6bu; v ^= v >> 13; v *= 0xc2b2ae35u; v ^= v >> 16;
        return v;
    };
    return mix((uint32_t)seed) ^ mix((uint32_t)(seed >> 32) + 0x9e3779b9u * (idx + 1u));
}

static size_t image_elems(rfv_handle h) { return (size_t)h->cfg.in_channels * h->cfg.image_size * h->cfg.image_size; }

RFV_EXPORT int rfv_velocity(rfv_handle h, const float* x, const float* t, float* v, int64_t batch, void* stream) {
    if (!h || !x || !t || !v || batch < 1) return fail(RFV_ERR_INVALID, "bad argument");
    if (h->cfg.in_channels != h->cfg.out_channels) return fail(RFV_ERR_INVALID, "in_channels != out_channels");
    const size_t ie = image_elems(h);
    RFV_TRY(h->enter((cudaStream_t)stream));
    for (int64_t b0 = 0; b0 < batch; b0 += h->cap) {
        RunCtx rc;
        rc.B = (int)std::min<int64_t>(h->cap, batch - b0);
        rc.x = x + b0 * ie; rc.t = t + b0; rc.out = v + b0 * ie; rc.mode = 0;
        RFV_TRY(h->run_forward(rc, (cudaStream_t)stream));
    }
    return h->leave((cudaStream_t)stream);
}

// One micro-batch's Euler loop as a resumable sequence, so that two lanes can be enqueued alternately from one thread.
struct EulerRun {
    rfv_engine* e = nullptr;
    cudaStream_t s = nullptr;
    float* x = nullptr;
    int B = 0, num_steps = 0;
    float* traj = nullptr;
    int save_every = 0;
    size_t traj_stride = 0;
    const float *tx0 = nullptr, *tx1 = nullptr;
    float* mse = nullptr;
    int i = 0;
    bool table = false;
};

static int euler_begin(EulerRun& r) {
    rfv_engine* h = r.e;
    const double dt = 1.0 / r.num_steps;  // Python double, like models/base_flow.py:158
    // every step's time is known now: one batched time-MLP + block-projection launch fills rows 0 .. num_steps-1 of the
    // projection table instead of two latency-bound single-row launches per step (0.07 ms of a 4.7 ms step at 256 images)
    r.table = r.num_steps > 1 && r.num_steps <= h->cap && !(h->cfg.flags & RFV_FLAG_TEMB_PER_STEP);
    r.i = 0;
    if (r.table) {
        fill_step_times_kernel<<<(r.num_steps + 255) / 256, 256, 0, r.s>>>(h->t_steps, r.num_steps, dt);
        CU_CHECK(cudaGetLastError());
        RunCtx rt;
        rt.B = r.num_steps; rt.t = h->t_steps; rt.temb_only = true;
        RFV_TRY(h->run_forward(rt, r.s));
    }
    return 0;
}

static int euler_step(EulerRun& r) {
    const double dt = 1.0 / r.num_steps;
    const int i = r.i++;
    RunCtx rc;
    rc.B = r.B; rc.x = r.x; rc.out = r.x; rc.mode = 1;
    rc.t = nullptr; rc.t_scalar = (float)(i * dt); rc.dt = (float)dt;
    if (r.table) rc.temb_row = i;
    if (r.traj && r.save_every > 0 && (i + 1) % r.save_every == 0) rc.traj = r.traj + (size_t)((i + 1) / r.save_every - 1) * r.traj_stride;
    if (r.mse) { rc.tgt_x0 = r.tx0; rc.tgt_x1 = r.tx1; rc.mse = r.mse + i; }
    return r.e->run_forward(rc, r.s);
}

static int euler_chunk(rfv_handle h, float* x, int B, int num_steps, float* traj, int save_every, size_t traj_stride,
                       const float* tx0, const float* tx1, float* mse, cudaStream_t s) {
    EulerRun r;
    r.e = h; r.s = s; r.x = x; r.B = B; r.num_steps = num_steps; r.traj = traj; r.save_every = save_every;
    r.traj_stride = traj_stride; r.tx0 = tx0; r.tx1 = tx1; r.mse = mse;
    RFV_TRY(euler_begin(r));
    while (r.i < num_steps) RFV_TRY(euler_step(r));
    return 0;
}

// The Euler loop of the micro-batch held in e->xbuf[buf] as one CUDA-graph launch on `s` (a non-default stream).  *done = false:
// the loop is not graph-able here (disabled, profiling, too many nodes, capture failed) and the caller enqueues it step by step.
static int euler_chunk_graph(rfv_engine* e, cudaStream_t s, int buf, int B, int num_steps, bool* done) {
    *done = false;
    if (!e->use_graphs || e->profiling || (size_t)num_steps * e->ops.size() > 16384) return 0;
    const auto key = std::make_tuple(B, num_steps, buf);
    auto it = e->loop_graphs.find(key);
    if (it == e->loop_graphs.end()) {
        if (e->loop_graphs.size() >= 32) return 0;
        rfv_engine::LoopGraph lg;
        const int64_t before = e->launches;
        cudaGraph_t gr = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return 0; }
        const int rc = euler_chunk(e, e->xbuf[buf], B, num_steps, nullptr, 0, 0, nullptr, nullptr, nullptr, s);
        const cudaError_t ce = cudaStreamEndCapture(s, &gr);
        lg.launches = e->launches - before;
        e->launches = before;
        if (rc == 0 && ce == cudaSuccess && gr && cudaGraphInstantiate(&lg.exec, gr, 0) != cudaSuccess) lg.exec = nullptr;
        if (gr) cudaGraphDestroy(gr);
        cudaGetLastError();
        if (rc != 0 || ce != cudaSuccess) lg.exec = nullptr;   // remembered: this shape is enqueued directly from now on
        it = e->loop_graphs.emplace(key, lg).first;
    }
    if (!it->second.exec) return 0;
    CU_CHECK(cudaGraphLaunch(it->second.exec, s));
    e->launches += it->second.launches;
    *done = true;
    return 0;
}

// the twin engine of the second lane: same configuration, weights copied device-to-device from this engine's fp32 masters
static int ensure_lane(rfv_handle h, cudaStream_t s) {
    if (h->lane) return 0;
    rfv_handle twin = nullptr;
    rfv_config cfg = h->cfg;
    cfg.flags |= RFV_FLAG_ONE_LANE;   // lanes do not nest
    RFV_TRY(rfv_create(&cfg, &twin));
    for (size_t i = 0; i < h->params.size(); ++i) {
        const Param& p = h->params[i];
        if (!p.loaded) continue;
        const int rc = rfv_set_tensor(twin, p.name.c_str(), p.f32, p.numel, s);
        if (rc) { rfv_destroy(twin); return rc; }
    }
    twin->profiling = false;
    h->lane = twin;
    return 0;
}

// Rows [0, n) in micro-batches, the first half of the chunks on lane 0 (this engine), the rest on lane 1 (the twin); the two
// chains advance one Euler step at a time, alternately, so both streams always hold queued work.  `begin(lane, engine, chunk
// index within the lane, first row, rows)` returns the device state buffer of the chunk (after any H2D copy it enqueues on the
// way); `end` is called when the chunk's last step is enqueued.
template <typename Begin, typename End>
static int run_two_lanes(rfv_handle h, int64_t n, int num_steps, cudaStream_t s0, cudaStream_t s1, Begin begin, End end) {
    const int64_t chunks = (n + h->cap - 1) / h->cap, half = (chunks + 1) / 2;
    struct Lane { rfv_engine* e; cudaStream_t s; int64_t c, c_end, k; bool open; EulerRun run; } ln[2] = {
        {h, s0, 0, half, 0, false, {}}, {h->lane, s1, half, chunks, 0, false, {}}};
    for (;;) {
        bool any = false;
        for (auto& l : ln) {
            if (l.c >= l.c_end) continue;
            any = true;
            if (!l.open) {
                const int64_t b0 = l.c * h->cap;
                const int B = (int)std::min<int64_t>(h->cap, n - b0);
                float* x = nullptr;
                RFV_TRY(begin((int)(&l - ln), l.e, l.k, b0, B, &x));
                // state in one of the engine's own buffers: the whole loop is one graph launch
                bool whole = false;
                if (x == l.e->xbuf[0] || x == l.e->xbuf[1]) RFV_TRY(euler_chunk_graph(l.e, l.s, x == l.e->xbuf[1] ? 1 : 0, B, num_steps, &whole));
                if (whole) {
                    RFV_TRY(end((int)(&l - ln), l.e, l.k, b0, B));
                    ++l.c; ++l.k;
                    continue;
                }
                l.run = EulerRun{};
                l.run.e = l.e; l.run.s = l.s; l.run.x = x; l.run.B = B; l.run.num_steps = num_steps;
                RFV_TRY(euler_begin(l.run));
                l.open = true;
            }
            RFV_TRY(euler_step(l.run));
            if (l.run.i >= num_steps) {
                RFV_TRY(end((int)(&l - ln), l.e, l.k, l.c * h->cap, l.run.B));
                l.open = false;
                ++l.c; ++l.k;
            }
        }
        if (!any) break;
    }
    return 0;
}

RFV_EXPORT int rfv_euler_sample(rfv_handle h, float* x, int64_t batch, int num_steps, float* traj, int save_every, void* stream) {
    if (!h || !x || batch < 1 || num_steps < 1) return fail(RFV_ERR_INVALID, "bad argument");
    const size_t ie = image_elems(h);
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));
    if (batch > h->cap && !traj && h->use_lanes && !h->profiling) {
        RFV_TRY(ensure_lane(h, s));
        // fork: both lane streams start after everything enqueued on the caller's stream (incl. weight uploads) ...
        CU_CHECK(cudaEventRecord(h->ev_fork, s));
        CU_CHECK(cudaStreamWaitEvent(h->s_cmp, h->ev_fork, 0));
        CU_CHECK(cudaStreamWaitEvent(h->lane->s_cmp, h->ev_fork, 0));
        CU_CHECK(cudaStreamWaitEvent(h->lane->s_cmp, h->lane->ev_weights, 0));
        // with the loop executor the chunk is staged through the engine's own state buffer (two device-to-device copies of
        // the chunk against hundreds of launches): captured graphs need fixed addresses
        const bool staged = h->use_graphs;
        RFV_TRY(run_two_lanes(h, batch, num_steps, h->s_cmp, h->lane->s_cmp,
            [&](int, rfv_engine* e, int64_t, int64_t b0, int B, float** xo) {
                if (!staged) { *xo = x + b0 * ie; return 0; }
                CU_CHECK(cudaMemcpyAsync(e->xbuf[0], x + b0 * ie, (size_t)B * ie * sizeof(float), cudaMemcpyDeviceToDevice, e->s_cmp));
                *xo = e->xbuf[0];
                return 0;
            },
            [&](int, rfv_engine* e, int64_t, int64_t b0, int B) {
                if (staged) CU_CHECK(cudaMemcpyAsync(x + b0 * ie, e->xbuf[0], (size_t)B * ie * sizeof(float), cudaMemcpyDeviceToDevice, e->s_cmp));
                return 0;
            }));
        // ... join: the caller's stream continues when both chains are done
        CU_CHECK(cudaEventRecord(h->ev_join[0], h->s_cmp));
        CU_CHECK(cudaEventRecord(h->ev_join[1], h->lane->s_cmp));
        CU_CHECK(cudaStreamWaitEvent(s, h->ev_join[0], 0));
        CU_CHECK(cudaStreamWaitEvent(s, h->ev_join[1], 0));
        return h->leave(s);
    }
    if (!traj && h->use_graphs && !h->profiling) {
        // one chain on the engine's compute stream (a captured graph cannot be recorded on the legacy default stream a caller
        // may hand in): fork, stage each chunk through xbuf[0], replay the loop graph, join
        CU_CHECK(cudaEventRecord(h->ev_fork, s));
        CU_CHECK(cudaStreamWaitEvent(h->s_cmp, h->ev_fork, 0));
        for (int64_t b0 = 0; b0 < batch; b0 += h->cap) {
            const int B = (int)std::min<int64_t>(h->cap, batch - b0);
            const size_t bytes = (size_t)B * ie * sizeof(float);
            CU_CHECK(cudaMemcpyAsync(h->xbuf[0], x + b0 * ie, bytes, cudaMemcpyDeviceToDevice, h->s_cmp));
            bool whole = false;
            RFV_TRY(euler_chunk_graph(h, h->s_cmp, 0, B, num_steps, &whole));
            if (!whole) RFV_TRY(euler_chunk(h, h->xbuf[0], B, num_steps, nullptr, 0, 0, nullptr, nullptr, nullptr, h->s_cmp));
            CU_CHECK(cudaMemcpyAsync(x + b0 * ie, h->xbuf[0], bytes, cudaMemcpyDeviceToDevice, h->s_cmp));
        }
        CU_CHECK(cudaEventRecord(h->ev_join[0], h->s_cmp));
        CU_CHECK(cudaStreamWaitEvent(s, h->ev_join[0], 0));
        return h->leave(s);
    }
    for (int64_t b0 = 0; b0 < batch; b0 += h->cap) {
        const int B = (int)std::min<int64_t>(h->cap, batch - b0);
        RFV_TRY(euler_chunk(h, x + b0 * ie, B, num_steps, traj ? traj + b0 * ie : nullptr, save_every, (size_t)batch * ie, nullptr,
                            nullptr, nullptr, s));
    }
    return h->leave(s);
}

RFV_EXPORT int rfv_euler_sample_host(rfv_handle h, const float* noise_host, float* out_host, int64_t n, int num_steps) {
    if (!h || !noise_host || !out_host || n < 1 || num_steps < 1) return fail(RFV_ERR_INVALID, "bad argument");
    const size_t ie = image_elems(h);
    CU_CHECK(cudaStreamWaitEvent(h->s_cmp, h->ev_weights, 0));  // uploads were enqueued on the caller's stream
    RFV_TRY(h->enter(h->s_cmp));
    // per chunk: H2D on the engine's copy-in stream -> Euler loop on its compute stream -> D2H on its copy-out stream, double
    // buffered (buffer k is free once its previous result has left the device)
    auto begin = [&](int, rfv_engine* e, int64_t k, int64_t b0, int B, float** xo) {
        const int kb = (int)(k & 1);
        if (k >= 2) CU_CHECK(cudaStreamWaitEvent(e->s_h2d, e->ev_out[kb], 0));
        CU_CHECK(cudaMemcpyAsync(e->xbuf[kb], noise_host + b0 * ie, (size_t)B * ie * sizeof(float), cudaMemcpyHostToDevice, e->s_h2d));
        CU_CHECK(cudaEventRecord(e->ev_in[kb], e->s_h2d));
        CU_CHECK(cudaStreamWaitEvent(e->s_cmp, e->ev_in[kb], 0));
        *xo = e->xbuf[kb];
        return 0;
    };
    auto end = [&](int, rfv_engine* e, int64_t k, int64_t b0, int B) {
        const int kb = (int)(k & 1);
        CU_CHECK(cudaEventRecord(e->ev_done[kb], e->s_cmp));
        CU_CHECK(cudaStreamWaitEvent(e->s_d2h, e->ev_done[kb], 0));
        CU_CHECK(cudaMemcpyAsync(out_host + b0 * ie, e->xbuf[kb], (size_t)B * ie * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        CU_CHECK(cudaEventRecord(e->ev_out[kb], e->s_d2h));
        return 0;
    };
    const bool lanes = n > h->cap && h->use_lanes && !h->profiling;
    if (lanes) {
        RFV_TRY(ensure_lane(h, h->s_cmp));
        CU_CHECK(cudaEventRecord(h->ev_fork, h->s_cmp));
        CU_CHECK(cudaStreamWaitEvent(h->lane->s_cmp, h->ev_fork, 0));   // a new twin's weights were written on this stream,
        CU_CHECK(cudaStreamWaitEvent(h->lane->s_cmp, h->lane->ev_weights, 0));   // later uploads on the caller's
        RFV_TRY(run_two_lanes(h, n, num_steps, h->s_cmp, h->lane->s_cmp, begin, end));
        CU_CHECK(cudaStreamSynchronize(h->lane->s_d2h));
        CU_CHECK(cudaStreamSynchronize(h->lane->s_cmp));
    } else {
        int64_t k = 0;
        for (int64_t b0 = 0; b0 < n; b0 += h->cap, ++k) {
            const int B = (int)std::min<int64_t>(h->cap, n - b0);
            float* xd = nullptr;
            RFV_TRY(begin(0, h, k, b0, B, &xd));
            bool whole = false;
            RFV_TRY(euler_chunk_graph(h, h->s_cmp, (int)(k & 1), B, num_steps, &whole));
            if (!whole) RFV_TRY(euler_chunk(h, xd, B, num_steps, nullptr, 0, 0, nullptr, nullptr, nullptr, h->s_cmp));
            RFV_TRY(end(0, h, k, b0, B));
        }
    }
    CU_CHECK(cudaStreamSynchronize(h->s_d2h));
    CU_CHECK(cudaStreamSynchronize(h->s_cmp));
    return h->leave(h->s_cmp);
}

RFV_EXPORT int rfv_straightness(rfv_handle h, const float* x0, const float* x1, int64_t batch, int num_points, float* dev_out, void* stream) {
    if (!h || !x0 || !x1 || !dev_out || batch < 1 || num_points < 1) return fail(RFV_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t ie = image_elems(h);
    RFV_TRY(h->enter(s));
    CU_CHECK(cudaMemsetAsync(dev_out, 0, (size_t)num_points * sizeof(float), s));
    for (int64_t b0 = 0; b0 < batch; b0 += h->cap) {
        const int B = (int)std::min<int64_t>(h->cap, batch - b0);
        CU_CHECK(cudaMemcpyAsync(h->scratch_x, x0 + b0 * ie, (size_t)B * ie * sizeof(float), cudaMemcpyDeviceToDevice, s));
        RFV_TRY(euler_chunk(h, h->scratch_x, B, num_points, nullptr, 0, 0, x0 + b0 * ie, x1 + b0 * ie, dev_out, s));
    }
    scale_kernel<<<(num_points + 255) / 256, 256, 0, s>>>(dev_out, num_points, 1.0f / (float)((double)batch * ie));
    CU_CHECK(cudaGetLastError());
    return h->leave(s);
}

RFV_EXPORT int rfv_fm_loss(rfv_handle h, const float* x0, const float* x1, const float* t, int64_t batch, float* loss_out, void* stream) {
    if (!h || !x0 || !x1 || !t || !loss_out || batch < 1) return fail(RFV_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t ie = image_elems(h);
    RFV_TRY(h->enter(s));
    CU_CHECK(cudaMemsetAsync(loss_out, 0, sizeof(float), s));
    for (int64_t b0 = 0; b0 < batch; b0 += h->cap) {
        RunCtx rc;
        rc.B = (int)std::min<int64_t>(h->cap, batch - b0);
        rc.x = x0 + b0 * ie; rc.x1 = x1 + b0 * ie; rc.t = t + b0; rc.mode = 2;
        rc.tgt_x0 = x0 + b0 * ie; rc.tgt_x1 = x1 + b0 * ie; rc.mse = loss_out;
        RFV_TRY(h->run_forward(rc, s));
    }
    scale_kernel<<<1, 32, 0, s>>>(loss_out, 1, 1.0f / (float)((double)batch * ie));
    CU_CHECK(cudaGetLastError());
    return h->leave(s);
}

// ---- training ---------------------------------------------------------------------------------------------
RFV_EXPORT int rfv_zero_grad(rfv_handle h, void* stream) {
    if (!h || !h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    RFV_TRY(h->enter((cudaStream_t)stream));
    CU_CHECK(cudaMemsetAsync(h->gflat, 0, (size_t)h->gtotal * sizeof(float), (cudaStream_t)stream));
    return h->leave((cudaStream_t)stream);
}

RFV_EXPORT int rfv_train_accumulate(rfv_handle h, const float* x0, const float* x1, const float* t, int64_t batch, float dropout_p,
                                    uint64_t seed, float* loss_out, void* stream) {
    if (!h || !x0 || !x1 || !t || !loss_out || batch < 1) return fail(RFV_ERR_INVALID, "bad argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail(RFV_ERR_INVALID, "dropout probability must be in [0,1)");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t ie = image_elems(h);
    RFV_TRY(h->enter(s));
    CU_CHECK(cudaMemsetAsync(loss_out, 0, sizeof(float), s));
    int64_t idx = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += h->cap, ++idx) {
        RunCtx rc;
        rc.B = (int)std::min<int64_t>(h->cap, batch - b0);
        rc.x = x0 + b0 * ie; rc.x1 = x1 + b0 * ie; rc.t = t + b0; rc.mode = 3;
        rc.tgt_x0 = rc.x; rc.tgt_x1 = rc.x1; rc.mse = loss_out;
        rc.out = h->dv_buf;
        rc.dt = (float)(2.0 / ((double)batch * (double)ie));   // d mean((v - target)^2) / dv
        rc.train = true;
        rc.drop_thresh = (uint32_t)std::lround((double)dropout_p * 65536.0);
        rc.drop_scale = rc.drop_thresh ? (float)(1.0 / (1.0 - (double)rc.drop_thresh / 65536.0)) : 1.f;
        rc.seed = mix_seed(seed, (uint32_t)idx);
        RFV_TRY(h->run_forward(rc, s));
        RFV_TRY(h->run_backward(rc, s, b0 + h->cap >= batch));   // the last chunk completes the gradients: bucket events
    }
    h->have_fwd = false;   // the kept activations no longer belong to an rfv_train_forward call
    scale_kernel<<<1, 32, 0, s>>>(loss_out, 1, 1.0f / (float)((double)batch * ie));
    CU_CHECK(cudaGetLastError());
    return h->leave(s);
}

RFV_EXPORT int rfv_train_forward(rfv_handle h, const float* x, const float* t, int64_t batch, float dropout_p, uint64_t seed,
                                 float* v_out, void* stream) {
    if (!h || !x || !t || !v_out || batch < 1) return fail(RFV_ERR_INVALID, "bad argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    if (batch > h->cap) return fail(RFV_ERR_STATE, "rfv_train_forward keeps the activations of ONE micro-batch: batch %lld > micro_batch %d", (long long)batch, h->cap);
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail(RFV_ERR_INVALID, "dropout probability must be in [0,1)");
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));
    RunCtx rc;
    rc.B = (int)batch;
    rc.x = x; rc.t = t; rc.out = v_out; rc.mode = 0;
    rc.train = true;
    rc.drop_thresh = (uint32_t)std::lround((double)dropout_p * 65536.0);
    rc.drop_scale = rc.drop_thresh ? (float)(1.0 / (1.0 - (double)rc.drop_thresh / 65536.0)) : 1.f;
    rc.seed = mix_seed(seed, 0u);
    h->have_fwd = false;
    RFV_TRY(h->run_forward(rc, s));
    h->fwd_rc = rc;
    h->have_fwd = true;
    return h->leave(s);
}

RFV_EXPORT int rfv_train_backward(rfv_handle h, const float* dv, int64_t batch, void* stream) {
    if (!h || !dv) return fail(RFV_ERR_INVALID, "null argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    if (!h->have_fwd) return fail(RFV_ERR_STATE, "rfv_train_backward without a preceding rfv_train_forward on this handle");
    if (batch != h->fwd_rc.B) return fail(RFV_ERR_INVALID, "rfv_train_backward: batch %lld, the forward call had %d", (long long)batch, h->fwd_rc.B);
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));
    CU_CHECK(cudaMemcpyAsync(h->dv_buf, dv, (size_t)batch * image_elems(h) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    h->have_fwd = false;   // the backward pass consumes (overwrites) the kept activations' gradient slots
    RFV_TRY(h->run_backward(h->fwd_rc, s));
    return h->leave(s);
}

RFV_EXPORT int rfv_reset_optimizer(rfv_handle h, void* stream) {
    if (!h) return fail(RFV_ERR_INVALID, "null handle");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));
    CU_CHECK(cudaMemsetAsync(h->mflat, 0, (size_t)h->gtotal * sizeof(float), s));
    CU_CHECK(cudaMemsetAsync(h->vflat, 0, (size_t)h->gtotal * sizeof(float), s));
    return h->leave(s);
}

RFV_EXPORT int rfv_grad_buffer(rfv_handle h, float** dev_ptr, int64_t* numel) {
    if (!h || !dev_ptr || !numel) return fail(RFV_ERR_INVALID, "null argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    *dev_ptr = h->gflat;
    *numel = h->gtotal;
    return 0;
}

RFV_EXPORT int rfv_grad_bucket_count(rfv_handle h) { return (h && h->train) ? (int)h->buckets.size() : 0; }

RFV_EXPORT int rfv_grad_bucket_info(rfv_handle h, int index, int64_t* offset, int64_t* numel) {
    if (!h || !h->train || index < 0 || index >= (int)h->buckets.size()) return fail(RFV_ERR_INVALID, "gradient bucket index out of range");
    if (offset) *offset = h->buckets[index].off;
    if (numel) *numel = h->buckets[index].numel;
    return 0;
}

RFV_EXPORT int rfv_grad_bucket_wait(rfv_handle h, int index, void* stream) {
    if (!h || !h->train || index < 0 || index >= (int)h->buckets.size()) return fail(RFV_ERR_INVALID, "gradient bucket index out of range");
    if (!h->buckets_valid) return fail(RFV_ERR_STATE, "no rfv_train_accumulate has recorded bucket events yet");
    for (auto& e : h->buckets[index].ev) CU_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, e, 0));
    return 0;
}

RFV_EXPORT int rfv_bind_param(rfv_handle h, const char* name, float* dev_ptr) {
    if (!h || !name) return fail(RFV_ERR_INVALID, "null argument");
    auto it = h->param_index.find(name);
    if (it == h->param_index.end()) return fail(RFV_ERR_INVALID, "unknown tensor '%s'", name);
    h->params[it->second].bound = dev_ptr;
    h->adam_dirty = true;
    return 0;
}

RFV_EXPORT int rfv_get_grad(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, float scale, void* stream) {
    if (!h || !name || !dev_ptr) return fail(RFV_ERR_INVALID, "null argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    auto it = h->param_index.find(name);
    if (it == h->param_index.end()) return fail(RFV_ERR_INVALID, "unknown tensor '%s'", name);
    const Param& p = h->params[it->second];
    if (p.numel != numel) return fail(RFV_ERR_INVALID, "tensor '%s': expected %lld elements", name, (long long)p.numel);
    RFV_TRY(h->enter((cudaStream_t)stream));
    unpack_grad_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(h->gflat, dev_ptr, make_seg(p), scale);
    CU_CHECK(cudaGetLastError());
    return h->leave((cudaStream_t)stream);
}

RFV_EXPORT int rfv_optimizer_step(rfv_handle h, const rfv_adamw* hp, float* grad_norm_out, void* stream) {
    if (!h || !hp) return fail(RFV_ERR_INVALID, "null argument");
    if (!h->train) return fail(RFV_ERR_STATE, "engine was not created with RFV_FLAG_TRAIN");
    if (hp->step < 1) return fail(RFV_ERR_INVALID, "step must be >= 1 (bias correction)");
    cudaStream_t s = (cudaStream_t)stream;
    RFV_TRY(h->enter(s));
    if (h->adam_dirty) {
        std::vector<AdamSeg> segs;
        for (auto& p : h->params) segs.push_back(make_seg(p));
        CU_CHECK(cudaMemcpyAsync(h->d_segs, segs.data(), segs.size() * sizeof(AdamSeg), cudaMemcpyHostToDevice, s));
        CU_CHECK(cudaStreamSynchronize(s));  // `segs` is a host temporary
        h->adam_dirty = false;
    }
    sumsq_kernel<<<SUMSQ_BLOCKS, 256, 0, s>>>(h->gflat, (size_t)h->gtotal, hp->grad_scale, h->norm_partial);
    sumsq_final_kernel<<<1, 256, 0, s>>>(h->norm_partial, SUMSQ_BLOCKS, h->norm2);
    CU_CHECK(cudaGetLastError());
    AdamHyper a;
    a.lr = hp->lr; a.beta1 = hp->beta1; a.beta2 = hp->beta2; a.eps = hp->eps; a.wd = hp->weight_decay;
    a.max_norm = hp->max_grad_norm; a.grad_scale = hp->grad_scale;
    a.bc1 = (float)(1.0 - std::pow((double)hp->beta1, (double)hp->step));
    a.bc2 = (float)(1.0 - std::pow((double)hp->beta2, (double)hp->step));
    adamw_kernel<<<h->n_adam_blocks, 256, 0, s>>>(h->d_segs, h->d_adam_blocks, h->gflat, h->mflat, h->vflat, h->norm2, a);
    CU_CHECK(cudaGetLastError());
    if (grad_norm_out) {
        CU_CHECK(cudaMemcpyAsync(grad_norm_out, h->norm2, sizeof(float), cudaMemcpyDeviceToDevice, s));
        sqrt_kernel<<<1, 32, 0, s>>>(grad_norm_out);
        CU_CHECK(cudaGetLastError());
    }
    // refresh every packed / derived copy from the updated fp32 masters: ~350 small launches with fixed arguments,
    // captured once into a CUDA graph and replayed (0.7 ms of launch latency -> one graph launch)
    if (!h->repack_graph) {
        cudaStream_t cs = h->s_cmp;
        CU_CHECK(cudaStreamSynchronize(cs));
        CU_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (auto& p : h->params)
            if (p.repack && rc == 0) rc = p.repack(cs);
        cudaGraph_t gr = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cs, &gr);
        if (rc != 0) { if (gr) cudaGraphDestroy(gr); return rc; }
        if (ce != cudaSuccess) return fail(RFV_ERR_CUDA, "repack graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&h->repack_graph, gr, 0);
        cudaGraphDestroy(gr);
        if (ce != cudaSuccess) return fail(RFV_ERR_CUDA, "repack graph instantiate failed: %s", cudaGetErrorString(ce));
    }
    CU_CHECK(cudaGraphLaunch(h->repack_graph, s));
    CU_CHECK(cudaEventRecord(h->ev_weights, s));
    return h->leave(s);
}

RFV_EXPORT int64_t rfv_launch_count(rfv_handle h, int reset) {
    if (!h) return 0;
    int64_t v = h->launches;
    if (reset) h->launches = 0;
    if (h->lane) {
        v += h->lane->launches;
        if (reset) h->lane->launches = 0;
    }
    return v;
}

RFV_EXPORT double rfv_flops_per_image(rfv_handle h) { return h ? h->flops_per_image : 0.0; }

RFV_EXPORT int64_t rfv_debug_activation(rfv_handle h, const char* name, float* dev_out, int64_t capacity, void* stream) {
    if (!h || !name || !dev_out) return fail(RFV_ERR_INVALID, "null argument");
    if (!h->keep_acts) return fail(RFV_ERR_STATE, "engine was created without the keep-activations flag (4)");
    auto it = h->named_acts.find(name);
    if (it == h->named_acts.end()) return fail(RFV_ERR_INVALID, "unknown activation '%s'", name);
    const Act& a = *it->second;
    const int64_t per_img = (int64_t)a.C * a.H * a.W;
    const int B = (int)std::min<int64_t>(h->cap, capacity / per_img);
    if (B < 1) return fail(RFV_ERR_INVALID, "capacity too small");
    nhwc_to_nchw_kernel<<<512, 256, 0, (cudaStream_t)stream>>>(a.p, dev_out, B, a.C, a.H * a.W);
    if (cudaGetLastError() != cudaSuccess) return fail(RFV_ERR_CUDA, "debug copy launch failed");
    return (int64_t)B * per_img;
}

// ---------------------------------------------------------------------------------------------------------
// quality metrics on device tensors (metrics.cuh; utils/metrics.py:39-116 of the reference) -- stateless entry points
// ---------------------------------------------------------------------------------------------------------
RFV_EXPORT int rfv_metrics_mean(const float* x, int64_t n, int64_t d, double* mu, void* stream) {
    if (!x || !mu || n < 1 || d < 1 || n > (1 << 30) || d > (1 << 30)) return fail(RFV_ERR_INVALID, "rfv_metrics_mean: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(mu, 0, (size_t)d * sizeof(double), s));
    const int slices = (int)std::min<int64_t>(64, (n + 31) / 32);
    metrics_colsum_kernel<<<dim3((unsigned)((d + 255) / 256), slices), 256, 0, s>>>(x, (int)n, (int)d, mu);
    metrics_scale_kernel<<<(unsigned)((d + 255) / 256), 256, 0, s>>>(mu, (int)d, 1.0 / (double)n);
    CU_CHECK(cudaGetLastError());
    return 0;
}

RFV_EXPORT int rfv_metrics_covariance(const float* x, const double* mu, int64_t n, int64_t d, double* sigma, void* stream) {
    if (!x || !mu || !sigma || d < 1 || d > (1 << 20)) return fail(RFV_ERR_INVALID, "rfv_metrics_covariance: bad argument");
    if (n < 2 || n > (1 << 30)) return fail(RFV_ERR_INVALID, "rfv_metrics_covariance: needs at least two samples (got %lld)", (long long)n);
    const unsigned t = (unsigned)((d + MG_T - 1) / MG_T);
    metrics_gram_kernel<true><<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(x, x, mu, mu, (int)d, (int)d, (int)n, (int)d, 1.0 / (double)(n - 1), sigma);
    CU_CHECK(cudaGetLastError());
    return 0;
}

RFV_EXPORT int rfv_metrics_fid_terms(const float* x1, const double* mu1, int64_t n1, const float* x2, const double* mu2, int64_t n2, int64_t d,
                                     double* gram, double* terms, void* stream) {
    if (!x1 || !mu1 || !x2 || !mu2 || !gram || !terms || d < 1 || d > (1 << 30)) return fail(RFV_ERR_INVALID, "rfv_metrics_fid_terms: bad argument");
    if (n1 < 2 || n2 < 2 || n1 > (1 << 20) || n2 > (1 << 20))
        return fail(RFV_ERR_INVALID, "rfv_metrics_fid_terms: needs at least two samples per set (got %lld, %lld)", (long long)n1, (long long)n2);
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(terms, 0, 3 * sizeof(double), s));
    metrics_sqdiff_kernel<<<64, 256, 0, s>>>(mu1, mu2, (int)d, terms);
    metrics_sqdev_kernel<<<1024, 256, 0, s>>>(x1, mu1, (int)n1, (int)d, 1.0 / (double)(n1 - 1), terms + 1);
    metrics_sqdev_kernel<<<1024, 256, 0, s>>>(x2, mu2, (int)n2, (int)d, 1.0 / (double)(n2 - 1), terms + 2);
    const double scale = 1.0 / std::sqrt((double)(n1 - 1) * (double)(n2 - 1));
    metrics_gram_kernel<false><<<dim3((unsigned)((n2 + MG_T - 1) / MG_T), (unsigned)((n1 + MG_T - 1) / MG_T)), 256, 0, s>>>(
        x1, x2, mu1, mu2, (int)n1, (int)n2, (int)d, (int)d, scale, gram);
    CU_CHECK(cudaGetLastError());
    return 0;
}

RFV_EXPORT int rfv_metrics_ssim(const float* a, const float* b, int64_t batch, int channels, int height, int width, float data_range,
                                double* out, void* stream) {
    if (!a || !b || !out || batch < 1 || batch > 65535 || channels < 1 || channels > 65535)
        return fail(RFV_ERR_INVALID, "rfv_metrics_ssim: bad argument");
    if (height < SS_WIN || width < SS_WIN)
        return fail(RFV_ERR_INVALID, "rfv_metrics_ssim: the 7x7 window exceeds the image extent (%d x %d)", height, width);
    const size_t smem = (size_t)2 * (SS_ROWS + 2 * SS_PAD) * width * sizeof(float);
    if (smem > 48 * 1024) return fail(RFV_ERR_INVALID, "rfv_metrics_ssim: width %d unsupported (max %d)", width, (int)(48 * 1024 / (8 * (SS_ROWS + 2 * SS_PAD))));
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(out, 0, (size_t)batch * sizeof(double), s));
    const int hout = height - 2 * SS_PAD, wout = width - 2 * SS_PAD;
    const double c1 = (0.01 * (double)data_range) * (0.01 * (double)data_range), c2 = (0.03 * (double)data_range) * (0.03 * (double)data_range);
    metrics_ssim_kernel<<<dim3((unsigned)((hout + SS_ROWS - 1) / SS_ROWS), (unsigned)channels, (unsigned)batch), 256, smem, s>>>(
        a, b, channels, height, width, c1, c2, 1.0 / ((double)hout * wout * channels), out);
    CU_CHECK(cudaGetLastError());
    return 0;
}

RFV_EXPORT int rfv_set_profiling(rfv_handle h, int enabled) {
    if (!h) return fail(RFV_ERR_INVALID, "null handle");
    h->profiling = enabled != 0;
    if (enabled) h->prof.clear();
    return 0;
}

RFV_EXPORT int rfv_profile_report(rfv_handle h, char* buf, int buf_len) {
    if (!h || !buf || buf_len < 1) return fail(RFV_ERR_INVALID, "bad argument");
    std::string out;
    char line[512];
    std::map<std::string, double> flops, bytes;
    for (auto& op : h->ops) { flops[op.kind + " " + op.label] = op.flops; bytes[op.kind + " " + op.label] = op.bytes; }
    for (auto& blk : h->bwd_blocks)
        for (auto& op : blk) { flops[op.kind + " " + op.label] = op.flops; bytes[op.kind + " " + op.label] = op.bytes; }
    // "<kind> <label>\t<total ms>\t<launches>\t<algorithmic FLOPs per image>\t<algorithmic HBM bytes per image>"
    for (auto& kv : h->prof) {
        snprintf(line, sizeof(line), "%s\t%.6f\t%lld\t%.1f\t%.1f\n", kv.first.c_str(), kv.second.first, (long long)kv.second.second,
                 flops[kv.first], bytes[kv.first]);
        out += line;
    }
    snprintf(buf, buf_len, "%s", out.c_str());
    return 0;
}
