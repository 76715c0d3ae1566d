// Kernel-emulation harness -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
//
// Compiles the library's kernel sources with g++ against the fiber shim and exposes the fused gap-evaluation
//   predict_layers_kernel -> [reach_caps_kernel] -> fast_pull_kernel<FastDescProv, DESC, WRAP, MAXT, HINT>
// as one C function, with the launch shape chosen by the caller.  tests/test_kernel_emulation_cpu.py compares its outputs
// with the CPU oracle / the C model of the fast kernel.
#include "cuda_emu.h"

#include "../../rl_mpc_lanemerging_b200/csrc/mpc_common.cuh"

static char g_err[256];
int mpc_set_error(int code, const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }
int mpc_set_cuda_error(cudaError_t, const char *what) { snprintf(g_err, sizeof(g_err), "%s", what); return MPC_E_CUDA; }

extern "C" long long emu_node_count() { return emu::S().nodes; }   // nodes finalised by the fast kernel since emu_plan began

#include "../../rl_mpc_lanemerging_b200/csrc/mpc_derive.h"
#include "../../rl_mpc_lanemerging_b200/csrc/mpc_predict.cu"
#undef FULL
#include "../../rl_mpc_lanemerging_b200/csrc/mpc_reach.cu"
#include "../../rl_mpc_lanemerging_b200/csrc/mpc_fast.cu"

#include <vector>

extern "C" const char *emu_last_error() { return g_err; }
extern "C" void emu_set_preempt(unsigned long long seed) { emu::set_preempt(seed); }
extern "C" long emu_cas_lost() { return emu::S().cas_lost; }
extern "C" void emu_ring_traffic(long long *out2) { out2[0] = emu::S().cell_reads; out2[1] = emu::S().cas_issued; }

// One fused gap-evaluation of B states on the emulated device.
//   threads: block size of the fast kernel (multiple of 32, <= 1024); ring: label ring capacity in cells (0 = full row, no wrap)
//   hint_cost (nullable) / hint_scale / hint_retry: the cost-hint ladder; use_caps: reachability pruning (hinted solves)
//   use_bound: 0 = unbounded pass, 1 = standard zone bound
// Outputs as mpc_plan; fallback[b] = 1 when the kernel handed the problem back (ring overflow / saturation).
extern "C" int emu_plan(const mpc_params *params, int B, int nmax, const double *ego, const double *cars_x, const double *cars_v,
                        const int32_t *n_cars, int threads, int ring, int use_bound, const double *hint_cost, double hint_scale,
                        double hint_retry, int use_caps, int32_t *idx, double *s_seq, double *cost, int32_t *reached,
                        uint8_t *crash, double *min_dist, double *start_s, uint8_t *fallback, int32_t *num_t_out) {
    DevParams P;
    emu::S().nodes = 0;
    int rc = derive_params(params, &P);
    if (rc) return rc;
    if (!P.fast_ok) return mpc_set_error(MPC_E_INVALID, "fast mode not available for these params");
    const int T = P.num_t;
    if (num_t_out) *num_t_out = T;
    std::vector<LayerDesc> desc((size_t)B * T);
    std::vector<double> s0(B), ds(B);
    std::vector<int32_t> num_s(B);
    // K1a
    emu::launch((B + 3) / 4, 128, 0, [&] { predict_layers_kernel<false>(P, B, ego, cars_x, cars_v, n_cars, nmax, desc.data(), s0.data(), ds.data(), num_s.data(), nullptr, nullptr); });
    if (start_s) for (int b = 0; b < B; b++) start_s[b] = s0[b];
    // reachability caps
    const int stride = ((P.num_s_max + 63) / 64 + 7) & ~7;
    std::vector<unsigned short> capb;
    const bool hinted = hint_cost != nullptr;
    if (hinted && use_caps && P.zone_ok && P.vstar_c > 0 && use_bound) {
        capb.assign((size_t)B * T * stride, 0);
        emu::launch((B + RC_WARPS - 1) / RC_WARPS, 32 * RC_WARPS, 0, [&] { reach_caps_kernel(P, B, desc.data(), num_s.data(), capb.data(), stride); });
    }
    // K3
    const int W = (P.num_s_max + 15) & ~15;
    const bool wrap = ring > 0 && ring < W;
    const int Wc = wrap ? (ring & ~7) : W;
    const size_t clamp_bytes = (4 * (((size_t)P.num_s_max + 31) / 32) + 4) * 4 + 16;
    const size_t smem = (size_t)Wc * 16 + clamp_bytes;
    std::vector<uint16_t> bp((size_t)T * W);
    std::vector<int> counters(16, 0);
    std::vector<int32_t> fb_list(2 * (size_t)B + 2, -1);
    SolveIO io; memset(&io, 0, sizeof(io));
    io.ego = ego;
    io.idx = idx; io.s_seq = s_seq; io.cost = cost; io.reached = reached; io.crash = crash; io.min_dist = min_dist;
    io.bp = bp.data(); io.bp_stride = W;
    io.work_counter = &counters[0]; io.fallback_list = fb_list.data(); io.fallback_count = &counters[2];
    io.hint_cost = hint_cost; io.hint_scale = hint_scale; io.hint_retry = hint_retry;
    io.capb = capb.empty() ? nullptr : capb.data(); io.cap_stride = stride;
    const unsigned long long bound = use_bound ? P.bound_fx : ~0ULL;
    const LayerDesc *d = desc.data();
#define EMU_RUN(WRAPV, MAXTV, HINTV) \
    emu::launch(1, threads, smem, [&] { fast_pull_kernel<FastDescProv, true, WRAPV, MAXTV, HINTV>(P, B, io, d, nullptr, nullptr, 0, Wc, bound); })
#define EMU_SHAPE(MAXTV) \
    do { if (wrap) { if (hinted) EMU_RUN(true, MAXTV, true); else EMU_RUN(true, MAXTV, false); } \
         else { if (hinted) EMU_RUN(false, MAXTV, true); else EMU_RUN(false, MAXTV, false); } } while (0)
    if (threads <= 192) EMU_SHAPE(192);
    else if (threads <= 512) EMU_SHAPE(512);
    else EMU_SHAPE(1024);
    if (fallback) {
        memset(fallback, 0, B);
        for (int i = 0; i < counters[2]; i++) if (fb_list[i] >= 0 && fb_list[i] < B) fallback[fb_list[i]] = 1;
    }
    return MPC_OK;
}

// The fast mode's chain of mpc_plan (run_solve in mpc_api.cu): fast32_kernel on all B states, then fast_pull_kernel (64-bit words) on the
// problems it handed on.  threads32 / ring32: launch shape of the 32-bit-key kernel (ring32 = 0: full row); threads / ring: of the other.
// handed[b]: bit 0 = handed on by fast32_kernel, bit 1 = with the "bounded attempt failed" flag, bit 2 = handed back by fast_pull_kernel.
// info3 = {f32_frac, f32_bound, f32_ok}.
extern "C" int emu_plan32(const mpc_params *params, int B, int nmax, const double *ego, const double *cars_x, const double *cars_v,
                          const int32_t *n_cars, int threads32, int ring32, int threads, int ring, int32_t *idx, double *s_seq,
                          double *cost, int32_t *reached, uint8_t *crash, double *min_dist, uint8_t *handed, int64_t *info3) {
    DevParams P;
    emu::S().nodes = 0;
    int rc = derive_params(params, &P);
    if (rc) return rc;
    if (info3) { info3[0] = P.f32_frac; info3[1] = P.f32_bound; info3[2] = P.f32_ok; }
    if (!P.fast_ok || !P.f32_ok) return mpc_set_error(MPC_E_INVALID, "32-bit-key kernel not available for these params");
    if (B <= 0) return MPC_OK;                   // (info3 only)
    const int T = P.num_t;
    std::vector<LayerDesc> desc((size_t)B * T);
    std::vector<double> s0(B), ds(B);
    std::vector<int32_t> num_s(B);
    emu::launch((B + 3) / 4, 128, 0, [&] { predict_layers_kernel<false>(P, B, ego, cars_x, cars_v, n_cars, nmax, desc.data(), s0.data(), ds.data(), num_s.data(), nullptr, nullptr); });
    const int W = (P.num_s_max + 15) & ~15;
    const size_t clamp_bytes = (4 * (((size_t)P.num_s_max + 31) / 32) + 4) * 4 + 16;
    std::vector<uint16_t> bp((size_t)T * W);
    std::vector<int> counters(16, 0);
    std::vector<int32_t> fb_list(3 * (size_t)B + 3, -1);
    SolveIO io; memset(&io, 0, sizeof(io));
    io.ego = ego;
    io.idx = idx; io.s_seq = s_seq; io.cost = cost; io.reached = reached; io.crash = crash; io.min_dist = min_dist;
    io.bp = bp.data(); io.bp_stride = W;
    const LayerDesc *d = desc.data();
    {   // first attempt: 32-bit keys
        const bool wrap = ring32 > 0 && ring32 < W;
        const int Wc = wrap ? (ring32 & ~7) : W;
        if (Wc < 1024 || 2 * Wc < W) return mpc_set_error(MPC_E_INVALID, "ring32 too small");
        const size_t smem = (size_t)Wc * 12 + fast32_smem_head(P.num_s_max);
        io.work_counter = &counters[0]; io.fallback_list = fb_list.data() + 2 * (size_t)B; io.fallback_count = &counters[5];
#define EMU_RUN32(WRAPV, MAXTV) emu::launch(1, threads32, smem, [&] { fast32_kernel<FastDescProv, WRAPV, MAXTV>(P, B, io, d, nullptr, nullptr, 0, Wc); })
        if (threads32 <= 192) { if (wrap) EMU_RUN32(true, 192); else EMU_RUN32(false, 192); }
        else if (threads32 <= 512) { if (wrap) EMU_RUN32(true, 512); else EMU_RUN32(false, 512); }
        else { if (wrap) EMU_RUN32(true, 1024); else EMU_RUN32(false, 1024); }
    }
    if (handed) {
        memset(handed, 0, B);
        for (int i = 0; i < counters[5]; i++) { const int32_t e = fb_list[2 * (size_t)B + i]; handed[e & 0x3fffffff] = (uint8_t)(1 | ((e >> 30) & 1) << 1); }
    }
    if (counters[5] > 0) {   // the problems it handed on: 64-bit words
        const bool wrap = ring > 0 && ring < W;
        const int Wc = wrap ? (ring & ~7) : W;
        const size_t smem = (size_t)Wc * 16 + clamp_bytes;
        io.work_counter = &counters[6]; io.subset = fb_list.data() + 2 * (size_t)B; io.B_dev = &counters[5];
        io.fallback_list = fb_list.data(); io.fallback_count = &counters[2];
        const unsigned long long bound = P.bound_fx;
#define EMU_RUN64(WRAPV, MAXTV) emu::launch(1, threads, smem, [&] { fast_pull_kernel<FastDescProv, true, WRAPV, MAXTV, false>(P, B, io, d, nullptr, nullptr, 0, Wc, bound); })
        if (threads <= 192) { if (wrap) EMU_RUN64(true, 192); else EMU_RUN64(false, 192); }
        else if (threads <= 512) { if (wrap) EMU_RUN64(true, 512); else EMU_RUN64(false, 512); }
        else { if (wrap) EMU_RUN64(true, 1024); else EMU_RUN64(false, 1024); }
        if (handed) for (int i = 0; i < counters[2]; i++) if (fb_list[i] >= 0 && fb_list[i] < B) handed[fb_list[i]] |= 4;
    }
    return MPC_OK;
}

// HighwayState.predict_step_without_ego through the K4 kernel
extern "C" int emu_predict_step_without_ego(const mpc_params *params, int B, int nmax, const double *ego, const double *cars_x,
                                            const double *cars_v, const double *cars_a, const int32_t *n_cars, double dt, double mcd,
                                            double *ego_out, double *ox, double *ov, double *oa, uint8_t *crashed) {
    DevParams P;
    int rc = derive_params(params, &P);
    if (rc) return rc;
    emu::launch((B + 3) / 4, 128, 0, [&] { predict_step_without_ego_kernel(P, B, nmax, ego, cars_x, cars_v, cars_a, n_cars, dt, mcd, ego_out, ox, ov, oa, crashed); });
    return MPC_OK;
}
