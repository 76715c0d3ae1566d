// C-ABI entry points of libmpcb200 (see include/mpcb200.h).  Host-side only: parameter derivation,
// scratch ownership, launch configuration.  There is deliberately no CPU fallback: without a CUDA
// device every compute entry point fails with MPC_E_NODEVICE.
#include "mpc_common.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// ---- launchers implemented in the kernel translation units ----------------------------------------
cudaError_t launch_exact_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st);
cudaError_t launch_exact_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                               int dist_f32, int stride, cudaStream_t st);
cudaError_t launch_fast_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st);
cudaError_t launch_fast_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                              int dist_f32, int stride, cudaStream_t st);
cudaError_t launch_check_sorted(const DevParams &P, int B, const LayerDesc *desc, const double *s0, const double *ds,
                                const int32_t *ns, unsigned long long *mismatches, cudaStream_t st);
cudaError_t launch_finer_fit(const DevParams &P, int B, int T, const double *s_seq, const int32_t *reached, const double *ego,
                             int max_iter, double tol, double *fine, int fine_stride, int32_t *n_fine, double *speed,
                             int32_t *iters, cudaStream_t st, const int32_t *subset = nullptr, const int *count = nullptr);
int qp_max_fine();
int exact_occupancy(int threads, size_t smem);
int fast_occupancy(int threads, size_t smem, int wrap);
cudaError_t launch_fast32_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st);
cudaError_t launch_fast32_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                                int dist_f32, int stride, cudaStream_t st);
int fast32_occupancy(int threads, size_t smem, int wrap);
size_t fast32_head_bytes(int num_s_max);
cudaError_t launch_predict_layers(const DevParams &P, int B, int nmax, const double *ego, const double *cx, const double *cv,
                                  const int32_t *n, LayerDesc *desc, double *s0, double *ds, int32_t *ns, cudaStream_t st, const int32_t *subset = nullptr, const int *count = nullptr);
cudaError_t launch_rasterise(const DevParams &P, int B, int stride_s, const LayerDesc *desc, const double *s0, const double *ds,
                             const int32_t *ns, uint8_t *obstacles, void *distances, int dist_f32, cudaStream_t st);
cudaError_t launch_predict_step(const DevParams &P, int B, int nmax, const double *ego, const double *cx, const double *cv,
                                const double *ca, const int32_t *n, const double *sel, double dt, double mcd, double *ego_out,
                                double *ox, double *ov, double *oa, uint8_t *crashed, cudaStream_t st);
cudaError_t launch_state_vector(const DevParams &P, int B, int nmax, const double *ego, const double *cx, const double *cv,
                                const double *ca, const int32_t *n, float *out, int stride, cudaStream_t st);
cudaError_t launch_speed_from_jerk(const DevParams &P, int B, const double *ego, const double *jerk, double *speed, cudaStream_t st);
cudaError_t launch_predict_step_without_ego(const DevParams &P, int B, int nmax, const double *ego, const double *cx,
                                            const double *cv, const double *ca, const int32_t *n, double dt, double mcd,
                                            double *ego_out, double *ox, double *ov, double *oa, uint8_t *crashed, cudaStream_t st);
cudaError_t launch_env_step(const DevParams &P, const mpc_env_params &E, int B, int nmax, double *ego, double *cx, double *cv, double *ca,
                            int32_t *n, double *prev_acc, double *delay, int32_t *ticks, const double *jerk, const double *u_spawn,
                            const double *gap_u, const double *first_u, const double *speed_z, const double *delay_u, double *reward,
                            uint8_t *flags, double *proj_jerk, cudaStream_t st);
cudaError_t launch_compact_mask(const uint8_t *mask, int B, int32_t *subset, int *count, cudaStream_t st);
cudaError_t launch_reach_caps(const DevParams &P, int B, const LayerDesc *desc, const int32_t *num_s, unsigned short *capb,
                              int stride, cudaStream_t st);
cudaError_t launch_rollout_step(const DevParams &P, int B, int nmax, double *ego, double *cx, double *cv, double *ca,
                                const int32_t *n, const double *jerk, double dt, double mcd, double stop_x, int step,
                                uint8_t *alive, double *sel_speed, double *roll_s, int roll_stride, int32_t *roll_len,
                                uint8_t *crash_pred, cudaStream_t st);

// ---- errors ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int mpc_set_error(int code, const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }
int mpc_set_cuda_error(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return MPC_E_CUDA;
}
extern "C" const char *mpc_last_error(void) { return g_err; }
#ifdef MPC_HOST_EMU      // tests/emu only: nodes the fast kernel has finalised so far
extern "C" long long mpc_emu_nodes(void) { return emu::S().nodes; }
#endif
extern "C" int mpc_abi_version(void) { return MPC_ABI_VERSION; }
extern "C" int mpc_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

extern "C" void mpc_default_params(mpc_params *p) {
    p->s_disc = 0.05; p->t_disc = 0.30; p->future_s = 150.0; p->future_t = 5.0;
    p->start_uncertainty = 0.0; p->uncertainty_per_second = 0.0;
    p->d_weight = 10.0; p->v_weight = 0.5; p->a_weight = 10.0; p->j_weight = 10.0;
    p->desired_speed = 30.0; p->max_speed = 30.0;
    p->a_min = -6.0; p->a_max = 4.5; p->j_min = -5.0; p->j_max = 5.0;
    p->min_allowed_distance = 5.0; p->crash_min_s = 20.0; p->car_length = 5.0;
    p->max_predicted_decel = -4.0; p->tick_length = 0.2; p->sensor_radius = 125.0;
    p->combination_min_distance = 5.1;
}

// ---- handle ------------------------------------------------------------------------------------------
struct mpc_handle {
    DevParams P;
    int device, max_batch, nmax;
    int sm_count; size_t smem_optin;
    // launch configuration
    int W;                       // full-row label-array length (cells), also the back-pointer row stride
    size_t smem;                 // dynamic shared memory of the exact kernel (0 -> it uses global label scratch)
    int threads, grid_exact;
    int Wc, wrap_fast, threads_fast, grid_fast; size_t smem_fast;   // fast kernel: ring capacity / launch shape
    struct Shape64 { int Wc, wrap, threads, grid; size_t smem; };   // a launch shape of the 64-bit kernel (grid = 0: not available)
    Shape64 hand_few, hand_many;                                    // hand-over list: few large blocks (short lists) / many smaller ones (long lists)
    int alt_mode;                                                   // MPC_HANDOVER_ALT: 0 main shape only, 1 by the list's length (default), 2 always hand_many (tests)
    size_t smem_fast_big;        // full-row variant used to re-solve ring overflows (0 = does not fit)
    int Wc32, wrap32, threads32, grid32; size_t smem32;             // 32-bit-key kernel (mpc_fast32.cuh): ring capacity / launch shape; grid32 == 0: not used
    int Wc32b, wrap32b, threads32b, grid32b; size_t smem32b;        // its second shape (wider ring) for the problems that outgrew the first; grid32b == 0: none
    int nb32, nb32b, adapt32, threads32_override, calls32, last_B32; // blocks per SM of the two shapes; calls left in which shape A may still be widened
    int use_bound;               // first fast pass runs under DevParams::bound_fx (MPC_FAST_BOUND=0|1 overrides)
    int grid_max;
    // scratch
    LayerDesc *desc; double *s0, *ds; int32_t *num_s;
    uint16_t *bp; int *counters; int32_t *fallback_list;
    uint16_t *bp_side;           // back-pointer scratch of the launch that runs on `side` (MPC_SIDE_BLOCKS blocks)
    unsigned long long *glab; unsigned *ghist;
    // staging for the host-buffer entry point
    double *st_ego, *st_cx, *st_cv, *st_ca; int32_t *st_n;
    int32_t *st_idx; double *st_seq, *st_cost, *st_mind, *st_s0; int32_t *st_reached; uint8_t *st_crash;
    unsigned short *capb; int cap_stride;      // reachability caps of hinted solves (allocated on first use)
    int32_t *mask_list; int *mask_count;        // masked calls: compacted episode list + its length (allocated on first use)
    int use_heur;                                // MPC_FAST_HEUR=0 disables the heuristic pruning of hinted solves (dev A/B)
    int probe_overlap;                           // MPC_PROBE_OVERLAP=1: mpc_plan_probed runs the probe plan on a second stream, next to the real predictor
    cudaStream_t aux; cudaEvent_t ev_fork, ev_join;
    int use_side; cudaStream_t side; cudaEvent_t ev_side_fork, ev_side_join;      // run_solve: fast32 shape B next to the 64-bit kernel's hand-over launch
    double hint_retry;                           // middle rung of the hinted ladder (MPC_HINT_RETRY, default 1.36 = 1.5 / 1.1; <= 1 disables)
    int64_t kernels_launched;
    // optional per-kernel timing (bench.py roofline): events around [predict | DP | fallback DP]
    int timing; cudaEvent_t ev[4]; int ev_valid;
};

#include "mpc_derive.h"

static void free_scratch(mpc_handle *h) {
    if (h->capb) { cudaFree(h->capb); h->capb = nullptr; }
    if (h->mask_list) { cudaFree(h->mask_list); h->mask_list = nullptr; }
    if (h->mask_count) { cudaFree(h->mask_count); h->mask_count = nullptr; }
    void *ptrs[] = {h->desc, h->s0, h->ds, h->num_s, h->bp, h->bp_side, h->counters, h->fallback_list, h->glab, h->ghist, h->st_ego,
                    h->st_cx, h->st_cv, h->st_ca, h->st_n, h->st_idx, h->st_seq, h->st_cost, h->st_mind, h->st_s0,
                    h->st_reached, h->st_crash};
    for (void *p : ptrs) if (p) cudaFree(p);
}

#define MPC_SIDE_BLOCKS 48          // grid of the launch on the side stream (fast32 wide-ring shape next to the 64-bit hand-over launch)

static int env_int(const char *name, int lo, int hi, int dflt) {
    const char *v = getenv(name);
    if (!v) return dflt;
    int x = atoi(v);
    if (x >= lo && x <= hi && (x % 32 == 0 || hi <= 1)) return x;
    fprintf(stderr, "libmpcb200: ignoring %s=%s (expected a multiple of 32 in [%d, %d])\n", name, v, lo, hi);   // dev overrides only
    return dflt;
}

static int configure(mpc_handle *h, int want_nb32 = 0) {
    const DevParams &P = h->P;
    h->W = (P.num_s_max + 15) & ~15;             // rows of the dense grids / back-pointer scratch: multiples of 16 cells (vector and bulk stores)
    const size_t static_smem = 11264;                // static shared of the kernels (upper bound: 10.2 KB in the fast kernel) + 1 KB/block reserve
    // ---- exact kernel: 24 B per cell, full row ----
    size_t need = (size_t)h->W * 24;
    if (need + static_smem <= h->smem_optin) {
        h->smem = need;
        int bps = (int)(h->smem_optin / (need + static_smem));
        h->threads = env_int("MPC_EXACT_THREADS", 64, 512, bps >= 3 ? 256 : 512);
        int occ = exact_occupancy(h->threads, h->smem);
        h->grid_exact = h->sm_count * (occ < 1 ? 1 : occ);
    } else {                                         // e.g. H = 100: labels live in per-block global scratch
        h->smem = 0;
        h->threads = 512;
        h->grid_exact = h->sm_count * 2;
    }
    // ---- fast kernel: 16 B per cell (one packed 64-bit word, double buffered) behind a ring window ----
    const size_t clamp_bytes = (4 * (((size_t)P.num_s_max + 31) / 32) + 4) * 4 + 16;   // behind the word buffers: two clamp bit arrays, two blocked-cell bit arrays (+2 words each)
    // Ring sized for two blocks per SM when that still covers 2/3 of the row (frontier spans measured: <= 0.6 of the row on
    // ordinary traffic; H=50: ring = 0.73 row, 7.7+0.5 ms vs 9.7 ms with one block per SM), else for one block per SM (long horizons).  MPC_FAST_BLOCKS=32|64 overrides (1 | 2 blocks/SM).
    // MPC_FAST_BLOCKS = 32 * (blocks per SM, 1..4) and MPC_FAST_THREADS override the shape (dev sweeps).
    // Blocks per SM: as many (<= 5) as still leave a ring of 2/3 of the row -- the barrier per layer makes a block latency bound,
    // more resident blocks hide it (H=17: 5 x 192 threads, H=25: 3 x 384, H=50: 2 x 512, H=100: 1 x 1024).
    int auto_blocks = 1;
    for (int nb = 2; nb <= 5; nb++) {
        const size_t per = (h->smem_optin + 1024) / nb;
        if (per > 1024 + static_smem + clamp_bytes && ((per - 1024 - static_smem - clamp_bytes) / 16) * 3 >= (size_t)h->W * 2) auto_blocks = nb;
    }
    int fast_blocks = env_int("MPC_FAST_BLOCKS", 32, 160, 32 * auto_blocks) / 32;
    size_t cap = ((h->smem_optin + 1024) / fast_blocks - 1024 - static_smem - clamp_bytes) / 16;
    h->wrap_fast = (size_t)h->W > cap;
    h->Wc = h->wrap_fast ? (int)(cap & ~(size_t)7) : h->W;
    h->smem_fast = (size_t)h->Wc * 16 + clamp_bytes;
    int bps = (int)(h->smem_optin / (h->smem_fast + static_smem));
    h->threads_fast = env_int("MPC_FAST_THREADS", 64, 1024, bps >= 5 ? 192 : (bps >= 4 ? 256 : (bps >= 3 ? 384 : (bps >= 2 ? 512 : 1024))));
    int occ = fast_occupancy(h->threads_fast, h->smem_fast, h->wrap_fast);
    h->grid_fast = P.fast_ok ? h->sm_count * (occ < 1 ? 1 : occ) : 0;
    h->smem_fast_big = ((size_t)h->W * 16 + clamp_bytes + static_smem <= h->smem_optin) ? (size_t)h->W * 16 + clamp_bytes : 0;
    // Launch shapes for the hand-over list of the 32-bit-key kernel.  Every problem of that list takes 0.1-0.3 ms whatever its size
    // (latency: two barriers per layer), so a SHORT list wants few large blocks (lowest latency per problem: H=17, 311 entries:
    // 2 x 512 threads 0.11 ms against 0.13 ms for the bulk shape 5 x 192) and a LONG list must not leave a slow problem for a second
    // round: beyond ~9/8 of the resident blocks the wider grid wins although each block is a little slower (H=50: 327 entries 0.38
    // vs 0.43 ms, 339 / 382 entries 0.52 / 0.53 vs 0.43 / 0.44 ms).  Both launches are enqueued; the list's length, known on the
    // device only, lets one of them work (SolveIO::n_lo / n_hi).
    auto shape64 = [&](int nb, mpc_handle::Shape64 *S) {
        S->grid = 0;
        const size_t per = (h->smem_optin + 1024) / nb;
        if (!P.fast_ok || per <= 1024 + static_smem + clamp_bytes) return;
        const size_t c = (per - 1024 - static_smem - clamp_bytes) / 16;
        if (c * 5 < (size_t)h->W * 2) return;                                        // ring >= 0.4 of the row (what outgrows it goes to the full-row launch)
        S->wrap = (size_t)h->W > c;
        S->Wc = S->wrap ? (int)(c & ~(size_t)7) : h->W;
        S->smem = (size_t)S->Wc * 16 + clamp_bytes;
        S->threads = nb >= 5 ? 192 : (nb >= 4 ? 256 : (nb >= 3 ? 384 : (nb >= 2 ? 512 : 1024)));
        const int o = fast_occupancy(S->threads, S->smem, S->wrap);
        if (o >= nb) S->grid = h->sm_count * nb;
    };
    h->alt_mode = env_int("MPC_HANDOVER_ALT", 0, 2, 1);
    h->hand_few.grid = 0; h->hand_many.grid = 0;
    if (h->alt_mode) {
        shape64(fast_blocks < 2 ? fast_blocks : 2, &h->hand_few);
        shape64(fast_blocks > 3 ? fast_blocks : 3, &h->hand_many);
        if (h->hand_few.grid == 0 || h->hand_many.grid == 0 || h->hand_few.grid >= h->hand_many.grid) { h->hand_few.grid = 0; h->hand_many.grid = 0; }
    }
    h->use_bound = env_int("MPC_FAST_BOUND", 0, 1, 1) && P.bound_fx != 0;
    // ---- 32-bit-key kernel: 12 B per cell (three rotating arrays of 32-bit words) behind a ring window ----
    // Two launch shapes.  A: as many blocks per SM (<= 5) as leave a ring of half the row -- frontier spans at H=50: <= 0.46 of the
    // row in moderate traffic, but 8-12 % of the problems in low / default / fast traffic need 0.6-0.9 of it.  B: the shape with the
    // fewest blocks whose ring covers 0.88 of the row; it takes the problems whose frontier outgrew ring A (a device-side list, like
    // every other hand-over).  MPC_F32_BLOCKS = 32 * blocks per SM and MPC_F32_THREADS override shape A (dev sweeps); MPC_FAST32=0
    // switches the kernel off (the 64-bit kernel does everything).
    h->grid32 = 0; h->grid32b = 0;
    { const char *e = getenv("MPC_FAST32");
      if (P.fast_ok && P.f32_ok && h->use_bound && !(e && e[0] == '0')) {
        // dynamic shared memory of fast32_kernel: head (its tables, staging, bit arrays: ~18 KB at H=50) + 12 B per ring cell; the only
        // static shared memory left is the back-track row of finish_problem (0.6 KB), plus the 1 KB per-block reserve
        const size_t head32 = fast32_head_bytes(P.num_s_max), static32 = 1024 + 1024;
        auto ring_of = [&](int nb) -> size_t {
            const size_t per = (h->smem_optin + 1024) / nb;
            return per > static32 + head32 ? (per - static32 - head32) / 12 : 0;
        };
        auto shape = [&](int nb, int threads_override, int *Wc, int *wrap, size_t *smem, int *threads, int *grid) {
            const size_t cap = ring_of(nb);
            *grid = 0;
            *wrap = (size_t)h->W > cap;
            *Wc = *wrap ? (int)(cap & ~(size_t)7) : h->W;
            if (*Wc < 1024 || 2 * *Wc < h->W) return;        // prologue scratch row; ring() folds an index once
            *smem = (size_t)*Wc * 12 + head32;
            const int bps = (int)(h->smem_optin / (*smem + static32 - 1024));
            *threads = threads_override ? threads_override : (bps >= 5 ? 192 : (bps >= 4 ? 256 : (bps >= 3 ? 384 : (bps >= 2 ? 512 : 1024))));
            const int occ = fast32_occupancy(*threads, *smem, *wrap);
            *grid = h->sm_count * (occ < 1 ? 1 : occ);
        };
        int autoA = 1, autoB = 1;
        for (int nb = 2; nb <= 5; nb++) {
            if (ring_of(nb) * 2 >= (size_t)h->W) autoA = nb;
            if (ring_of(nb) * 100 >= (size_t)h->W * 88) autoB = nb;
        }
        const int forced = env_int("MPC_F32_BLOCKS", 32, 192, 0) / 32;
        if (want_nb32 > 0) autoA = want_nb32;                // (re-configuration after an adaptive step, see run_solve)
        h->nb32 = forced ? forced : autoA; h->nb32b = autoB;
        h->threads32_override = env_int("MPC_F32_THREADS", 64, 1024, 0);
        if (want_nb32 == 0) h->adapt32 = forced ? 0 : 4;
        shape(h->nb32, h->threads32_override, &h->Wc32, &h->wrap32, &h->smem32, &h->threads32, &h->grid32);
        if (h->grid32 > 0 && h->wrap32 && autoB < h->nb32) {
            shape(autoB, 0, &h->Wc32b, &h->wrap32b, &h->smem32b, &h->threads32b, &h->grid32b);
            if (h->Wc32b <= h->Wc32) h->grid32b = 0;
        }
      } }
    { const char *e = getenv("MPC_SIDE_STREAM"); h->use_side = !(e && e[0] == '0'); }
    { const char *e = getenv("MPC_FAST_HEUR"); h->use_heur = !(e && e[0] == '0'); }
    { const char *e = getenv("MPC_PROBE_OVERLAP"); h->probe_overlap = (e && e[0] == '1'); }
    { const char *e = getenv("MPC_HINT_RETRY"); h->hint_retry = e ? atof(e) : 1.36; if (!(h->hint_retry >= 0.0 && h->hint_retry < 100.0)) h->hint_retry = 1.36; }
    h->grid_max = h->grid_fast > h->grid_exact ? h->grid_fast : h->grid_exact;
    if (h->grid32 > h->grid_max) h->grid_max = h->grid32;
    if (h->grid32b > h->grid_max) h->grid_max = h->grid32b;
    if (h->hand_many.grid > h->grid_max) h->grid_max = h->hand_many.grid;
    if (h->hand_few.grid > h->grid_max) h->grid_max = h->hand_few.grid;
    return MPC_OK;
}

static int alloc_scratch(mpc_handle *h) {
    const DevParams &P = h->P;
    size_t B = (size_t)h->max_batch, T = (size_t)P.num_t, N = (size_t)h->nmax;
    MPC_CUDA_OK(cudaMalloc(&h->desc, B * T * sizeof(LayerDesc)));
    MPC_CUDA_OK(cudaMalloc(&h->s0, B * 8)); MPC_CUDA_OK(cudaMalloc(&h->ds, B * 8)); MPC_CUDA_OK(cudaMalloc(&h->num_s, B * 4));
    MPC_CUDA_OK(cudaMalloc(&h->bp, (size_t)h->grid_max * T * h->W * sizeof(uint16_t)));
    MPC_CUDA_OK(cudaMalloc(&h->bp_side, (size_t)MPC_SIDE_BLOCKS * T * h->W * sizeof(uint16_t)));
    MPC_CUDA_OK(cudaMalloc(&h->counters, 16 * sizeof(int)));
    MPC_CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    MPC_CUDA_OK(cudaEventCreateWithFlags(&h->ev_side_fork, cudaEventDisableTiming));
    MPC_CUDA_OK(cudaEventCreateWithFlags(&h->ev_side_join, cudaEventDisableTiming));
    MPC_CUDA_OK(cudaMalloc(&h->fallback_list, 4 * B * 4));
    if (h->smem == 0) {
        MPC_CUDA_OK(cudaMalloc(&h->glab, (size_t)h->grid_exact * 2 * h->W * 8));
        MPC_CUDA_OK(cudaMalloc(&h->ghist, (size_t)h->grid_exact * 2 * h->W * 4));
    }
    MPC_CUDA_OK(cudaMalloc(&h->st_ego, B * 4 * 8));
    MPC_CUDA_OK(cudaMalloc(&h->st_cx, B * N * 8)); MPC_CUDA_OK(cudaMalloc(&h->st_cv, B * N * 8)); MPC_CUDA_OK(cudaMalloc(&h->st_ca, B * N * 8));
    MPC_CUDA_OK(cudaMalloc(&h->st_n, B * 4));
    MPC_CUDA_OK(cudaMalloc(&h->st_idx, B * T * 4)); MPC_CUDA_OK(cudaMalloc(&h->st_seq, B * T * 8));
    MPC_CUDA_OK(cudaMalloc(&h->st_cost, B * 8)); MPC_CUDA_OK(cudaMalloc(&h->st_mind, B * 8)); MPC_CUDA_OK(cudaMalloc(&h->st_s0, B * 8));
    MPC_CUDA_OK(cudaMalloc(&h->st_reached, B * 4)); MPC_CUDA_OK(cudaMalloc(&h->st_crash, B));
    return MPC_OK;
}

extern "C" int mpc_create(const mpc_params *p, int device, int max_batch, int nmax, mpc_handle **out) {
    if (!p || !out || max_batch <= 0) return mpc_set_error(MPC_E_INVALID, "mpc_create: bad argument");
    if (nmax <= 0 || nmax > MPC_NMAX) return mpc_set_error(MPC_E_CAPACITY, "mpc_create: nmax must be in [1,32]");
    int ndev = mpc_device_count();
    if (ndev <= 0) return mpc_set_error(MPC_E_NODEVICE, "no CUDA device: libmpcb200 has no CPU fallback");
    if (device < 0 || device >= ndev) return mpc_set_error(MPC_E_INVALID, "mpc_create: bad device index");
    MPC_CUDA_OK(cudaSetDevice(device));
    mpc_handle *h = (mpc_handle *)calloc(1, sizeof(mpc_handle));
    int rc = derive_params(p, &h->P);
    if (rc) { free(h); return rc; }
    h->device = device; h->max_batch = max_batch; h->nmax = nmax;
    cudaDeviceProp prop;
    MPC_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    configure(h);
    rc = alloc_scratch(h);
    if (rc) {
        if (h->side) cudaStreamDestroy(h->side);
        if (h->ev_side_fork) cudaEventDestroy(h->ev_side_fork);
        if (h->ev_side_join) cudaEventDestroy(h->ev_side_join);
        free_scratch(h); free(h); return rc;
    }
    *out = h;
    return MPC_OK;
}

extern "C" int mpc_set_params(mpc_handle *h, const mpc_params *p) {
    if (!h || !p) return mpc_set_error(MPC_E_INVALID, "mpc_set_params: bad argument");
    DevParams D;
    int rc = derive_params(p, &D);
    if (rc) return rc;
    if (D.num_t != h->P.num_t || D.num_s_max != h->P.num_s_max)
        return mpc_set_error(MPC_E_CAPACITY, "mpc_set_params: grid dimensions changed; create a new handle");
    h->P = D;
    return MPC_OK;
}

extern "C" int mpc_destroy(mpc_handle *h) {
    if (!h) return MPC_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->aux) { cudaStreamDestroy(h->aux); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    if (h->side) { cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_side_fork); cudaEventDestroy(h->ev_side_join); }
    free_scratch(h);
    free(h);
    return MPC_OK;
}

extern "C" int mpc_grid_dims(const mpc_handle *h, int *num_t, int *num_s_max) {
    if (!h) return mpc_set_error(MPC_E_INVALID, "null handle");
    if (num_t) *num_t = h->P.num_t;
    if (num_s_max) *num_s_max = h->P.num_s_max;
    return MPC_OK;
}

extern "C" int mpc_grid_stride(const mpc_handle *h) { return h ? h->W : 0; }

extern "C" int mpc_last_counters(const mpc_handle *h, int64_t *out2) {
    if (!h || !out2) return mpc_set_error(MPC_E_INVALID, "null argument");
    int c[16];
    MPC_CUDA_OK(cudaSetDevice(h->device));
    MPC_CUDA_OK(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    out2[0] = h->kernels_launched;
    out2[1] = c[2];          // problems the first fast launch handed back (c[4]: of those, the ones that needed the exact kernel)
    return MPC_OK;
}

extern "C" int mpc_fast32_info(const mpc_handle *h, int64_t *out6) {
    if (!h || !out6) return mpc_set_error(MPC_E_INVALID, "null argument");
    int c[16];
    MPC_CUDA_OK(cudaSetDevice(h->device));
    MPC_CUDA_OK(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    out6[0] = h->grid32 > 0; out6[1] = h->P.f32_frac; out6[2] = h->P.f32_bound; out6[3] = h->Wc32;
    // handed to the 64-bit kernel: with two shapes the flagged entries of the first (c[11]) + what the wide ring handed on (c[7])
    out6[4] = h->grid32 > 0 ? (h->grid32b > 0 ? c[11] + c[7] : c[5]) : 0; out6[5] = h->grid32 > 0 ? c[5] : 0;
    return MPC_OK;
}

extern "C" int mpc_set_timing(mpc_handle *h, int enable) {
    if (!h) return mpc_set_error(MPC_E_INVALID, "null handle");
    MPC_CUDA_OK(cudaSetDevice(h->device));
    if (enable && !h->ev[0]) for (int i = 0; i < 4; i++) MPC_CUDA_OK(cudaEventCreate(&h->ev[i]));
    h->timing = enable ? 1 : 0; h->ev_valid = 0;
    return MPC_OK;
}

// out3 = {traffic-predictor ms, DP kernel ms, fallback DP kernel ms} of the last mpc_plan / mpc_solve_dense call
// (waits for that call to finish).  Needs mpc_set_timing(h, 1) before the call.
extern "C" int mpc_last_kernel_ms(mpc_handle *h, float *out3) {
    if (!h || !out3) return mpc_set_error(MPC_E_INVALID, "null argument");
    if (!h->timing || !h->ev_valid) return mpc_set_error(MPC_E_INVALID, "timing not enabled or no timed call yet");
    MPC_CUDA_OK(cudaSetDevice(h->device));
    MPC_CUDA_OK(cudaEventSynchronize(h->ev[3]));
    out3[0] = 0.f;
    if (h->ev_valid == 1) MPC_CUDA_OK(cudaEventElapsedTime(&out3[0], h->ev[0], h->ev[1]));
    MPC_CUDA_OK(cudaEventElapsedTime(&out3[1], h->ev[1], h->ev[2]));
    MPC_CUDA_OK(cudaEventElapsedTime(&out3[2], h->ev[2], h->ev[3]));
    return MPC_OK;
}

static int check_batch(mpc_handle *h, int B) {
    if (!h) return mpc_set_error(MPC_E_INVALID, "null handle");
    if (B < 0) return mpc_set_error(MPC_E_INVALID, "negative batch");
    if (B > h->max_batch) return mpc_set_error(MPC_E_CAPACITY, "batch larger than the handle's max_batch");
    MPC_CUDA_OK(cudaSetDevice(h->device));
    return MPC_OK;
}

// ---- K1 ----------------------------------------------------------------------------------------------
extern "C" int mpc_build_grid(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                              const double *d_cars_a, const int32_t *d_n_cars, uint8_t *d_obstacles, void *d_distances,
                              int dist_f32, double *d_start_s, double *d_delta_s, int32_t *d_num_s, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_n_cars || !d_obstacles || !d_distances) return mpc_set_error(MPC_E_INVALID, "mpc_build_grid: null pointer");
    (void)d_cars_a;
    cudaStream_t st = (cudaStream_t)stream;
    h->ev_valid = 0;
    if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[0], st));
    MPC_CUDA_OK(launch_predict_layers(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_n_cars, h->desc, h->s0, h->ds, h->num_s, st));
    if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[1], st));
    MPC_CUDA_OK(launch_rasterise(h->P, B, h->W, h->desc, h->s0, h->ds, h->num_s, d_obstacles, d_distances, dist_f32, st));
    if (h->timing) { MPC_CUDA_OK(cudaEventRecord(h->ev[2], st)); MPC_CUDA_OK(cudaEventRecord(h->ev[3], st)); h->ev_valid = 1; }
    if (d_start_s) MPC_CUDA_OK(cudaMemcpyAsync(d_start_s, h->s0, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    if (d_delta_s) MPC_CUDA_OK(cudaMemcpyAsync(d_delta_s, h->ds, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    if (d_num_s) MPC_CUDA_OK(cudaMemcpyAsync(d_num_s, h->num_s, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    h->kernels_launched = 2;
    return MPC_OK;
}

// Self-test of the per-layer search structure the fast kernel uses: rebuilds the layer descriptors for the given
// states and compares, for every cell of every layer, the sorted O(1) lookup with the reference-order evaluation.
extern "C" int mpc_selftest_search(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                                   const int32_t *d_n_cars, int64_t *mismatching_cells, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_n_cars || !mismatching_cells) return mpc_set_error(MPC_E_INVALID, "mpc_selftest_search: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(h->counters + 8);
    MPC_CUDA_OK(cudaMemsetAsync(cnt, 0, 8, st));
    MPC_CUDA_OK(launch_predict_layers(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_n_cars, h->desc, h->s0, h->ds, h->num_s, st));
    MPC_CUDA_OK(launch_check_sorted(h->P, B, h->desc, h->s0, h->ds, h->num_s, cnt, st));
    unsigned long long v = 0;
    MPC_CUDA_OK(cudaMemcpyAsync(&v, cnt, 8, cudaMemcpyDeviceToHost, st));
    MPC_CUDA_OK(cudaStreamSynchronize(st));
    *mismatching_cells = (int64_t)v;
    h->kernels_launched = 2;
    return MPC_OK;
}

// ---- K2 / K3 -----------------------------------------------------------------------------------------
static int run_solve(mpc_handle *h, int B, int mode, bool dense, SolveIO io, const uint8_t *ob, const void *dist, int dist_f32,
                     int stride, cudaStream_t st) {
    if (mode != MPC_MODE_FAST && mode != MPC_MODE_EXACT) return mpc_set_error(MPC_E_INVALID, "unknown mode");
    if (mode == MPC_MODE_FAST && (!h->P.fast_ok || h->grid_fast == 0)) mode = MPC_MODE_EXACT;   // still on the GPU
    // Shape A adapts during the first calls on a handle: frontier widths depend on the traffic (H=50: <= 0.46 of the row in
    // moderate traffic, 0.9 for 10 % of the problems in low traffic), and a problem that outgrows ring A is solved twice.  When
    // more than 2 % of the previous call's problems did, A gets one block per SM less (a wider ring).  Results never depend on it.
    if (mode == MPC_MODE_FAST && !dense && h->grid32 > 0 && io.hint_cost == nullptr && h->adapt32 > 0 && h->calls32 > 0) {
        int c[16];
        MPC_CUDA_OK(cudaMemcpyAsync(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        MPC_CUDA_OK(cudaStreamSynchronize(st));
        h->adapt32--;
        if (c[9] * 50 > h->last_B32 && h->nb32 > h->nb32b) {
            const int keep = h->adapt32;
            configure(h, h->nb32 - 1);
            h->adapt32 = keep;
        }
    }
    MPC_CUDA_OK(cudaMemsetAsync(h->counters, 0, 16 * sizeof(int), st));
    io.bp = h->bp; io.bp_stride = h->W;
    io.fallback_list = h->fallback_list; io.fallback_count = h->counters + 2;
    // (io.subset / io.B_dev as passed: masked plans list the episodes to solve, everybody else passes NULL)
    SolveLaunch X;                                   // exact-kernel launch shape
    X.B = B; X.threads = h->threads; X.smem = h->smem; X.W = h->W; X.wrap = 0; X.bound = ~0ULL;
    X.glab = h->smem ? nullptr : h->glab; X.ghist = h->smem ? nullptr : h->ghist;
    X.grid = h->grid_exact < B ? h->grid_exact : B;
    cudaError_t e;
    if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[1], st));
    if (mode == MPC_MODE_EXACT) {
        io.work_counter = h->counters + 0;
        e = dense ? launch_exact_dense(h->P, X, io, ob, dist, dist_f32, stride, st) : launch_exact_desc(h->P, X, io, h->desc, st);
        if (e != cudaSuccess) return mpc_set_cuda_error(e, "exact solve launch");
        h->kernels_launched++;
        if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[2], st));
    } else {
        SolveLaunch F;
        F.B = B; F.threads = h->threads_fast; F.smem = h->smem_fast; F.W = h->Wc; F.wrap = h->wrap_fast;
        F.glab = nullptr; F.ghist = nullptr;
        F.bound = h->use_bound ? h->P.bound_fx : ~0ULL;
        F.grid = h->grid_fast < B ? h->grid_fast : B;
        if (h->grid32 > 0 && io.hint_cost == nullptr && (!dense || h->P.zone_ok)) {
            // First attempt by the 32-bit-key kernel (bounded, zones closed).  What it cannot finish -- plans that cross a penalty
            // zone or do not reach the horizon (list entries with bit 30: straight to the unbounded pass), frontiers wider than
            // its ring -- goes to the 64-bit kernel through a device-side list, like that kernel's own hand-backs below.
            h->calls32++; h->last_B32 = B;
            io.overflow_count = h->counters + 9; io.flagged_count = h->counters + 11;
            SolveLaunch F3 = F;
            F3.threads = h->threads32; F3.smem = h->smem32; F3.W = h->Wc32; F3.wrap = h->wrap32;
            F3.grid = h->grid32 < B ? h->grid32 : B;
            io.work_counter = h->counters + 0;
            io.fallback_list = h->fallback_list + 2 * (size_t)h->max_batch; io.fallback_count = h->counters + 5;
            e = dense ? launch_fast32_dense(h->P, F3, io, ob, dist, dist_f32, stride, st) : launch_fast32_desc(h->P, F3, io, h->desc, st);
            if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast32 solve launch");
            h->kernels_launched++;
            if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[2], st));
            const int32_t *handed = h->fallback_list + 2 * (size_t)h->max_batch; const int *handed_n = h->counters + 5;
            // Two kinds of entries in that list, two independent launches that run side by side: the flagged ones (bit 30: bounded
            // attempt failed) go to the 64-bit kernel's unbounded pass on the caller's stream; the frontiers that outgrew ring A
            // go once more through the 32-bit-key kernel with the wide ring, on a second stream (a single such problem occupies
            // one block for ~0.3 ms at H=50 -- serialised in front of the 64-bit launch it cost that much per step on the ranks
            // whose episodes held one).  What the wide ring cannot finish either follows in a second 64-bit launch after the join.
            const bool two_shapes = h->grid32b > 0;
            cudaStream_t sside = h->use_side ? h->side : st;        // MPC_SIDE_STREAM=0: one after the other on the caller's stream
            if (two_shapes) {
                if (h->use_side) {
                    MPC_CUDA_OK(cudaEventRecord(h->ev_side_fork, st));
                    MPC_CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_side_fork, 0));
                }
                SolveIO iob = io;
                F3.threads = h->threads32b; F3.smem = h->smem32b; F3.W = h->Wc32b; F3.wrap = h->wrap32b;
                F3.grid = h->grid32b < B ? h->grid32b : B;
                if (F3.grid > MPC_SIDE_BLOCKS) F3.grid = MPC_SIDE_BLOCKS;        // few problems by construction (shape A adapts above 2 %)
                iob.bp = h->bp_side;                                              // the two launches run at the same time: own scratch rows
                iob.work_counter = h->counters + 8;
                iob.subset = handed; iob.B_dev = handed_n; iob.skip_flagged = 1;
                iob.fallback_list = h->fallback_list + 3 * (size_t)h->max_batch; iob.fallback_count = h->counters + 7;
                iob.overflow_count = nullptr; iob.flagged_count = nullptr;
                e = dense ? launch_fast32_dense(h->P, F3, iob, ob, dist, dist_f32, stride, sside) : launch_fast32_desc(h->P, F3, iob, h->desc, sside);
                if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast32 wide-ring launch");
                h->kernels_launched++;
                if (h->use_side) MPC_CUDA_OK(cudaEventRecord(h->ev_side_join, h->side));
            }
            io.overflow_count = nullptr; io.flagged_count = nullptr;
            io.work_counter = h->counters + 6;
            io.subset = handed; io.B_dev = handed_n; io.only_flagged = two_shapes ? 1 : 0;
            io.fallback_list = h->fallback_list; io.fallback_count = h->counters + 2;
            // two shapes enqueued, the list's length (on the device) picks one: see configure()
            if (h->hand_many.grid > 0) {
                const int n_split = h->alt_mode == 2 ? -1 : h->hand_few.grid + h->hand_few.grid / 8;
                const mpc_handle::Shape64 *shapes[2] = {&h->hand_few, &h->hand_many};
                for (int v = 0; v < 2; v++) {
                    SolveLaunch FH = F;
                    FH.threads = shapes[v]->threads; FH.smem = shapes[v]->smem; FH.W = shapes[v]->Wc; FH.wrap = shapes[v]->wrap;
                    FH.grid = shapes[v]->grid < B ? shapes[v]->grid : B;
                    if (v == 0) { io.n_lo = 0; io.n_hi = n_split < 0 ? -1 : n_split; io.work_counter = h->counters + 6; }
                    else { io.n_lo = n_split + 1; io.n_hi = INT_MAX; io.work_counter = h->counters + 12; }
                    e = dense ? launch_fast_dense(h->P, FH, io, ob, dist, dist_f32, stride, st) : launch_fast_desc(h->P, FH, io, h->desc, st);
                    if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast solve launch (hand-backs of the 32-bit-key kernel)");
                    h->kernels_launched++;
                }
                io.n_lo = 0; io.n_hi = 0;
            } else {
                e = dense ? launch_fast_dense(h->P, F, io, ob, dist, dist_f32, stride, st) : launch_fast_desc(h->P, F, io, h->desc, st);
                if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast solve launch (hand-backs of the 32-bit-key kernel)");
                h->kernels_launched++;
            }
            io.only_flagged = 0;
            if (two_shapes) {
                if (h->use_side) MPC_CUDA_OK(cudaStreamWaitEvent(st, h->ev_side_join, 0));
                io.work_counter = h->counters + 10;
                io.subset = h->fallback_list + 3 * (size_t)h->max_batch; io.B_dev = h->counters + 7;
                e = dense ? launch_fast_dense(h->P, F, io, ob, dist, dist_f32, stride, st) : launch_fast_desc(h->P, F, io, h->desc, st);
                if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast solve launch (hand-backs of the wide ring)");
                h->kernels_launched++;
            }
        } else {
        io.work_counter = h->counters + 0;
        e = dense ? launch_fast_dense(h->P, F, io, ob, dist, dist_f32, stride, st) : launch_fast_desc(h->P, F, io, h->desc, st);
        if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast solve launch");
        h->kernels_launched++;
        if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[2], st));
        }
        // Problems the fast kernel handed back are re-solved on the device (their count is read on the device):
        // first by the fast kernel with a full-row window (ring overflow), then by the exact kernel (label saturation).
        // (A cost bound that proved too low is handled inside the kernel: the same block repeats the problem without it.)
        const int32_t *pending = h->fallback_list; const int *pending_n = h->counters + 2;
        if (h->wrap_fast && h->smem_fast_big) {
            SolveLaunch G = F;
            G.threads = 1024; G.smem = h->smem_fast_big; G.W = h->W; G.wrap = 0;      // one wide block per hard problem
            G.grid = h->sm_count < B ? h->sm_count : B; if (G.grid > 32) G.grid = 32;
            io.work_counter = h->counters + 3;
            io.subset = pending; io.B_dev = pending_n;
            io.fallback_list = h->fallback_list + h->max_batch; io.fallback_count = h->counters + 4;
            e = dense ? launch_fast_dense(h->P, G, io, ob, dist, dist_f32, stride, st) : launch_fast_desc(h->P, G, io, h->desc, st);
            if (e != cudaSuccess) return mpc_set_cuda_error(e, "fast full-row re-solve launch");
            h->kernels_launched++;
            pending = h->fallback_list + h->max_batch; pending_n = h->counters + 4;
        }
        X.grid = X.grid < 16 ? X.grid : 16;
        io.work_counter = h->counters + 1;
        io.subset = pending; io.B_dev = pending_n;
        io.fallback_list = nullptr; io.fallback_count = nullptr;
        e = dense ? launch_exact_dense(h->P, X, io, ob, dist, dist_f32, stride, st) : launch_exact_desc(h->P, X, io, h->desc, st);
        if (e != cudaSuccess) return mpc_set_cuda_error(e, "fallback solve launch");
        h->kernels_launched++;
    }
    if (h->timing) { MPC_CUDA_OK(cudaEventRecord(h->ev[3], st)); h->ev_valid = dense ? 2 : 1; }
    return MPC_OK;
}

extern "C" int mpc_solve_dense(mpc_handle *h, int B, int num_t, int num_s_stride, const uint8_t *d_obstacles, const void *d_distances,
                               int dist_f32, const double *d_start_s, const double *d_delta_s, const int32_t *d_num_s,
                               const double *d_v0, const double *d_a0, int mode, int32_t *d_idx, double *d_s_seq, double *d_cost,
                               int32_t *d_reached_t, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (num_t != h->P.num_t) return mpc_set_error(MPC_E_INVALID, "mpc_solve_dense: num_t does not match the handle's horizon");
    if (num_s_stride > h->W) return mpc_set_error(MPC_E_CAPACITY, "mpc_solve_dense: row stride larger than the handle's num_s_max");
    if (!d_obstacles || !d_distances || !d_start_s || !d_delta_s || !d_num_s || !d_v0 || !d_a0) return mpc_set_error(MPC_E_INVALID, "mpc_solve_dense: null pointer");
    SolveIO io; memset(&io, 0, sizeof(io));
    io.s0 = d_start_s; io.ds = d_delta_s; io.num_s = d_num_s; io.v0 = d_v0; io.a0 = d_a0;
    io.idx = d_idx; io.s_seq = d_s_seq; io.cost = d_cost; io.reached = d_reached_t;
    h->kernels_launched = 0;
    return run_solve(h, B, mode, true, io, d_obstacles, d_distances, dist_f32, num_s_stride, (cudaStream_t)stream);
}

// mpc_plan with an optional per-problem cost hint for the fast kernel (hint_cost == NULL: none)
// builds the compacted list of the episodes with mask[b] != 0 in h->mask_list / h->mask_count (no host round trip)
static int build_mask_list(mpc_handle *h, int B, const uint8_t *d_mask, cudaStream_t st) {
    if (!h->mask_list) {
        MPC_CUDA_OK(cudaMalloc(&h->mask_list, (size_t)h->max_batch * sizeof(int32_t)));
        MPC_CUDA_OK(cudaMalloc(&h->mask_count, sizeof(int)));
    }
    MPC_CUDA_OK(launch_compact_mask(d_mask, B, h->mask_list, h->mask_count, st));
    return MPC_OK;
}

static int plan_impl(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                     const int32_t *d_n_cars, int mode, const double *hint_cost, const int32_t *hint_reached, int hint_full_t,
                     double hint_scale, int32_t *d_idx, double *d_s_seq, double *d_cost, int32_t *d_reached_t,
                     uint8_t *d_crash, double *d_min_dist, double *d_start_s, cudaStream_t st, const char *who,
                     const uint8_t *d_mask = nullptr, cudaEvent_t wait_before_solve = nullptr) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_n_cars) return mpc_set_error(MPC_E_INVALID, who);
    h->ev_valid = 0;
    if (h->timing) MPC_CUDA_OK(cudaEventRecord(h->ev[0], st));
    h->kernels_launched = 0;
    const int32_t *subset = nullptr; const int *count = nullptr;
    if (d_mask) { rc = build_mask_list(h, B, d_mask, st); if (rc) return rc; subset = h->mask_list; count = h->mask_count; h->kernels_launched++; }
    MPC_CUDA_OK(launch_predict_layers(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_n_cars, h->desc, d_start_s ? d_start_s : h->s0, h->ds, h->num_s, st,
                                      subset, count));
    h->kernels_launched++;
    SolveIO io; memset(&io, 0, sizeof(io));
    io.subset = subset; io.B_dev = count;
    io.ego = d_ego;
    io.idx = d_idx; io.s_seq = d_s_seq; io.cost = d_cost; io.reached = d_reached_t; io.crash = d_crash; io.min_dist = d_min_dist;
    io.hint_cost = hint_cost; io.hint_reached = hint_reached; io.hint_full_t = hint_full_t; io.hint_scale = hint_scale;
    io.hint_retry = h->hint_retry;
    if (hint_cost && mode == MPC_MODE_FAST && h->P.fast_ok && h->P.zone_ok && h->P.vstar_c > 0 && h->use_bound && h->use_heur) {
        // reachability caps for the exact A*-style pruning of the hinted lean pass (mpc_reach.cu)
        if (!h->capb) {
            h->cap_stride = ((h->P.num_s_max + 63) / 64 + 7) & ~7;
            if (h->cap_stride > MPC_MAX_BUCKETS) h->cap_stride = MPC_MAX_BUCKETS;
            MPC_CUDA_OK(cudaMalloc(&h->capb, (size_t)h->max_batch * h->P.num_t * h->cap_stride * sizeof(unsigned short)));
        }
        MPC_CUDA_OK(launch_reach_caps(h->P, B, h->desc, h->num_s, h->capb, h->cap_stride, st));
        h->kernels_launched++;
        io.capb = h->capb; io.cap_stride = h->cap_stride;
    }
    if (wait_before_solve) MPC_CUDA_OK(cudaStreamWaitEvent(st, wait_before_solve, 0));     // (the hints come from another stream)
    return run_solve(h, B, mode, false, io, nullptr, nullptr, 0, 0, st);
}

extern "C" int mpc_plan(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v, const double *d_cars_a,
                        const int32_t *d_n_cars, int mode, int32_t *d_idx, double *d_s_seq, double *d_cost, int32_t *d_reached_t,
                        uint8_t *d_crash, double *d_min_dist, double *d_start_s, void *stream) {
    (void)d_cars_a;
    return plan_impl(h, B, d_ego, d_cars_x, d_cars_v, d_n_cars, mode, nullptr, nullptr, 0, 1.0, d_idx, d_s_seq, d_cost, d_reached_t,
                     d_crash, d_min_dist, d_start_s, (cudaStream_t)stream, "mpc_plan: null pointer");
}

// Masked plan: only the episodes with d_mask[b] != 0 are planned; the outputs of the others are left untouched.  The list of
// masked episodes is compacted on the device and its length is read there, so the call needs no host round trip -- what the
// combined controller's take-over (dqn.py:148-155: st.do_st_control for the episodes the planner vetoed) needs to stay asynchronous.
extern "C" int mpc_plan_masked(mpc_handle *h, int B, const uint8_t *d_mask, const double *d_ego, const double *d_cars_x,
                               const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars, int mode, int32_t *d_idx,
                               double *d_s_seq, double *d_cost, int32_t *d_reached_t, uint8_t *d_crash, double *d_min_dist,
                               double *d_start_s, void *stream) {
    (void)d_cars_a;
    if (!d_mask) return mpc_set_error(MPC_E_INVALID, "mpc_plan_masked: null mask");
    if (d_start_s) return mpc_set_error(MPC_E_INVALID, "mpc_plan_masked: start_s is not available for masked plans (pass NULL)");
    return plan_impl(h, B, d_ego, d_cars_x, d_cars_v, d_n_cars, mode, nullptr, nullptr, 0, 1.0, d_idx, d_s_seq, d_cost, d_reached_t, d_crash,
                     d_min_dist, nullptr, (cudaStream_t)stream, "mpc_plan_masked: null pointer", d_mask);
}

extern "C" int mpc_plan_hinted(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                               const double *d_cars_a, const int32_t *d_n_cars, int mode, const double *d_hint_cost,
                               const int32_t *d_hint_reached_t, int hint_full_t, double hint_scale, int32_t *d_idx, double *d_s_seq,
                               double *d_cost, int32_t *d_reached_t, uint8_t *d_crash, double *d_min_dist, double *d_start_s,
                               void *stream) {
    (void)d_cars_a;
    if (d_hint_cost && !(hint_scale > 0.0)) return mpc_set_error(MPC_E_INVALID, "mpc_plan_hinted: hint_scale must be positive");
    if (d_hint_cost && (d_hint_cost == d_cost || (d_hint_reached_t && d_hint_reached_t == d_reached_t)))
        return mpc_set_error(MPC_E_INVALID, "mpc_plan_hinted: the hint arrays must not alias the outputs");
    return plan_impl(h, B, d_ego, d_cars_x, d_cars_v, d_n_cars, mode, d_hint_cost, d_hint_reached_t, hint_full_t, hint_scale, d_idx,
                     d_s_seq, d_cost, d_reached_t, d_crash, d_min_dist, d_start_s, (cudaStream_t)stream, "mpc_plan_hinted: null pointer");
}

// Probe + plan: `probe` is a second handle on the same device whose Settings differ only in a coarser S/T_DISCRETIZATION
// (a grid ~100x smaller).  Its plan of the same states costs a few percent of the real one and its cost, scaled by the
// ratio of the step counts and by `margin`, is the first bound of the real solve.
extern "C" int mpc_plan_probed(mpc_handle *h, mpc_handle *probe, double margin, int B, const double *d_ego, const double *d_cars_x,
                               const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars, int32_t *d_idx,
                               double *d_s_seq, double *d_cost, int32_t *d_reached_t, uint8_t *d_crash, double *d_min_dist,
                               double *d_start_s, void *stream) {
    (void)d_cars_a;
    if (!h || !probe || h == probe) return mpc_set_error(MPC_E_INVALID, "mpc_plan_probed: two distinct handles are required");
    if (probe->device != h->device) return mpc_set_error(MPC_E_INVALID, "mpc_plan_probed: the handles live on different devices");
    if (B > probe->max_batch) return mpc_set_error(MPC_E_CAPACITY, "mpc_plan_probed: batch larger than the probe handle's max_batch");
    if (!(margin > 0.0) || probe->P.num_t < 2) return mpc_set_error(MPC_E_INVALID, "mpc_plan_probed: bad margin / probe horizon");
    cudaStream_t st = (cudaStream_t)stream, pst = st;
    cudaEvent_t join = nullptr;
    int rc = check_batch(h, B); if (rc) return rc;
    if (h->probe_overlap) {
        // the probe plan and the real predictor (+ reachability caps) are independent: run the probe on a second stream and join
        // before the real DP reads its costs.  The fork event also orders this probe after the previous call's DP (same scratch).
        if (!h->aux) {
            MPC_CUDA_OK(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
            MPC_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            MPC_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        MPC_CUDA_OK(cudaEventRecord(h->ev_fork, st));
        MPC_CUDA_OK(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
        pst = h->aux; join = h->ev_join;
    }
    rc = plan_impl(probe, B, d_ego, d_cars_x, d_cars_v, d_n_cars, MPC_MODE_FAST, nullptr, nullptr, 0, 1.0, nullptr, nullptr,
                   probe->st_cost, probe->st_reached, nullptr, nullptr, nullptr, pst, "mpc_plan_probed: null pointer");
    if (rc) return rc;
    if (join) MPC_CUDA_OK(cudaEventRecord(join, pst));
    const int64_t probe_launches = probe->kernels_launched;
    // every step costs about the same in both grids (the cost is a sum over steps of the same rates): scale by the step counts
    const double scale = margin * (double)(h->P.num_t - 1) / (double)(probe->P.num_t - 1);
    rc = plan_impl(h, B, d_ego, d_cars_x, d_cars_v, d_n_cars, MPC_MODE_FAST, probe->st_cost, probe->st_reached, probe->P.num_t - 1,
                   scale, d_idx, d_s_seq, d_cost, d_reached_t, d_crash, d_min_dist, d_start_s, st, "mpc_plan_probed: null pointer",
                   nullptr, join);
    if (rc == MPC_OK) h->kernels_launched += probe_launches;
    return rc;
}

// host-buffer variants: copy the state in, plan (plain or probed), copy the results out, synchronise
static int plan_host_impl(mpc_handle *h, mpc_handle *probe, double margin, int B, const double *h_ego, const double *h_cars_x,
                          const double *h_cars_v, const int32_t *h_n_cars, int mode, int32_t *h_idx, double *h_s_seq, double *h_cost,
                          int32_t *h_reached_t, uint8_t *h_crash, double *h_min_dist, double *h_start_s, void *stream, const char *who) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!h_ego || !h_cars_x || !h_cars_v || !h_n_cars) return mpc_set_error(MPC_E_INVALID, who);
    cudaStream_t st = (cudaStream_t)stream;
    size_t N = (size_t)h->nmax, T = (size_t)h->P.num_t, b = (size_t)B;
    MPC_CUDA_OK(cudaMemcpyAsync(h->st_ego, h_ego, b * 32, cudaMemcpyHostToDevice, st));
    MPC_CUDA_OK(cudaMemcpyAsync(h->st_cx, h_cars_x, b * N * 8, cudaMemcpyHostToDevice, st));
    MPC_CUDA_OK(cudaMemcpyAsync(h->st_cv, h_cars_v, b * N * 8, cudaMemcpyHostToDevice, st));
    MPC_CUDA_OK(cudaMemcpyAsync(h->st_n, h_n_cars, b * 4, cudaMemcpyHostToDevice, st));
    int32_t *d_idx = h_idx ? h->st_idx : nullptr; double *d_seq = h_s_seq ? h->st_seq : nullptr, *d_cost = h_cost ? h->st_cost : nullptr;
    int32_t *d_reached = h_reached_t ? h->st_reached : nullptr; uint8_t *d_crash = h_crash ? h->st_crash : nullptr;
    double *d_mind = h_min_dist ? h->st_mind : nullptr, *d_s0 = h_start_s ? h->st_s0 : nullptr;
    if (probe) rc = mpc_plan_probed(h, probe, margin, B, h->st_ego, h->st_cx, h->st_cv, nullptr, h->st_n, d_idx, d_seq, d_cost, d_reached,
                                    d_crash, d_mind, d_s0, stream);
    else rc = mpc_plan(h, B, h->st_ego, h->st_cx, h->st_cv, nullptr, h->st_n, mode, d_idx, d_seq, d_cost, d_reached, d_crash, d_mind, d_s0, stream);
    if (rc) return rc;
    if (h_idx) MPC_CUDA_OK(cudaMemcpyAsync(h_idx, h->st_idx, b * T * 4, cudaMemcpyDeviceToHost, st));
    if (h_s_seq) MPC_CUDA_OK(cudaMemcpyAsync(h_s_seq, h->st_seq, b * T * 8, cudaMemcpyDeviceToHost, st));
    if (h_cost) MPC_CUDA_OK(cudaMemcpyAsync(h_cost, h->st_cost, b * 8, cudaMemcpyDeviceToHost, st));
    if (h_reached_t) MPC_CUDA_OK(cudaMemcpyAsync(h_reached_t, h->st_reached, b * 4, cudaMemcpyDeviceToHost, st));
    if (h_crash) MPC_CUDA_OK(cudaMemcpyAsync(h_crash, h->st_crash, b, cudaMemcpyDeviceToHost, st));
    if (h_min_dist) MPC_CUDA_OK(cudaMemcpyAsync(h_min_dist, h->st_mind, b * 8, cudaMemcpyDeviceToHost, st));
    if (h_start_s) MPC_CUDA_OK(cudaMemcpyAsync(h_start_s, h->st_s0, b * 8, cudaMemcpyDeviceToHost, st));
    MPC_CUDA_OK(cudaStreamSynchronize(st));
    return MPC_OK;
}

extern "C" int mpc_plan_host(mpc_handle *h, int B, const double *h_ego, const double *h_cars_x, const double *h_cars_v, const double *h_cars_a,
                             const int32_t *h_n_cars, int mode, int32_t *h_idx, double *h_s_seq, double *h_cost, int32_t *h_reached_t,
                             uint8_t *h_crash, double *h_min_dist, double *h_start_s, void *stream) {
    (void)h_cars_a;
    return plan_host_impl(h, nullptr, 0.0, B, h_ego, h_cars_x, h_cars_v, h_n_cars, mode, h_idx, h_s_seq, h_cost, h_reached_t, h_crash,
                          h_min_dist, h_start_s, stream, "mpc_plan_host: null pointer");
}

extern "C" int mpc_plan_host_probed(mpc_handle *h, mpc_handle *probe, double margin, int B, const double *h_ego, const double *h_cars_x,
                                    const double *h_cars_v, const double *h_cars_a, const int32_t *h_n_cars, int32_t *h_idx,
                                    double *h_s_seq, double *h_cost, int32_t *h_reached_t, uint8_t *h_crash, double *h_min_dist,
                                    double *h_start_s, void *stream) {
    (void)h_cars_a;
    if (!probe) return mpc_set_error(MPC_E_INVALID, "mpc_plan_host_probed: null probe handle");
    return plan_host_impl(h, probe, margin, B, h_ego, h_cars_x, h_cars_v, h_n_cars, MPC_MODE_FAST, h_idx, h_s_seq, h_cost, h_reached_t,
                          h_crash, h_min_dist, h_start_s, stream, "mpc_plan_host_probed: null pointer");
}

// ---- finer_fit (st.py:584-723): tick-rate re-sampling + speed/accel/jerk projection of the plans ------------------------
extern "C" int mpc_finer_fit(mpc_handle *h, int B, const double *d_s_seq, const int32_t *d_reached_t, const double *d_ego,
                             double *d_fine, int fine_stride, int32_t *d_n_fine, double *d_speed, int32_t *d_iterations, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_s_seq || !d_reached_t || !d_ego || !d_fine || !d_n_fine || fine_stride < 2) return mpc_set_error(MPC_E_INVALID, "mpc_finer_fit: bad argument");
    if (!(h->P.p.tick_length > 0)) return mpc_set_error(MPC_E_INVALID, "mpc_finer_fit: tick_length must be positive");
    MPC_CUDA_OK(launch_finer_fit(h->P, B, h->P.num_t, d_s_seq, d_reached_t, d_ego, 40, 1e-9, d_fine, fine_stride, d_n_fine, d_speed,
                                 d_iterations, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

// mpc_finer_fit for the episodes with d_mask[b] != 0 only (rows of the others untouched); no host round trip
extern "C" int mpc_finer_fit_masked(mpc_handle *h, int B, const uint8_t *d_mask, const double *d_s_seq, const int32_t *d_reached_t,
                                    const double *d_ego, double *d_fine, int fine_stride, int32_t *d_n_fine, double *d_speed,
                                    int32_t *d_iterations, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_mask || !d_s_seq || !d_reached_t || !d_ego || !d_fine || !d_n_fine || fine_stride < 2) return mpc_set_error(MPC_E_INVALID, "mpc_finer_fit_masked: bad argument");
    if (!(h->P.p.tick_length > 0)) return mpc_set_error(MPC_E_INVALID, "mpc_finer_fit_masked: tick_length must be positive");
    rc = build_mask_list(h, B, d_mask, (cudaStream_t)stream); if (rc) return rc;
    MPC_CUDA_OK(launch_finer_fit(h->P, B, h->P.num_t, d_s_seq, d_reached_t, d_ego, 40, 1e-9, d_fine, fine_stride, d_n_fine, d_speed,
                                 d_iterations, (cudaStream_t)stream, h->mask_list, h->mask_count));
    h->kernels_launched = 2;
    return MPC_OK;
}

extern "C" int mpc_finer_fit_max_points(void) { return qp_max_fine(); }

// ---- K4 ----------------------------------------------------------------------------------------------
extern "C" int mpc_predict_step_with_ego(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                                         const double *d_cars_a, const int32_t *d_n_cars, const double *d_selected_speed, double dt,
                                         double min_crash_distance, double *d_ego_out, double *d_cars_x_out, double *d_cars_v_out,
                                         double *d_cars_a_out, uint8_t *d_crashed, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_n_cars || !d_selected_speed || !d_ego_out || !d_cars_x_out || !d_cars_v_out)
        return mpc_set_error(MPC_E_INVALID, "mpc_predict_step_with_ego: null pointer");
    MPC_CUDA_OK(launch_predict_step(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, d_selected_speed, dt,
                                    min_crash_distance, d_ego_out, d_cars_x_out, d_cars_v_out, d_cars_a_out, d_crashed, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

cudaError_t launch_krauss_step(const DevParams &P, int B, int nmax, double *ego, double *cx, double *cv, double *ca, const int32_t *n,
                               const double *sel, double dt, double mcd, double accel, double decel, double tau, double min_gap,
                               double max_speed, uint8_t *crashed, cudaStream_t st);
extern "C" int mpc_krauss_step(mpc_handle *h, int B, double *d_ego, double *d_cars_x, double *d_cars_v, double *d_cars_a,
                               const int32_t *d_n_cars, const double *d_selected_speed, double dt, double min_crash_distance,
                               double accel, double decel, double tau, double min_gap, double max_speed, uint8_t *d_crashed, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_cars_a || !d_n_cars || !d_selected_speed || !d_crashed)
        return mpc_set_error(MPC_E_INVALID, "mpc_krauss_step: null pointer");
    if (!(dt > 0) || !(decel > 0) || !(accel > 0) || !(tau >= 0)) return mpc_set_error(MPC_E_INVALID, "mpc_krauss_step: bad model parameter");
    MPC_CUDA_OK(launch_krauss_step(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, d_selected_speed, dt, min_crash_distance,
                                   accel, decel, tau, min_gap, max_speed, d_crashed, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

extern "C" int mpc_predict_step_without_ego(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                                            const double *d_cars_a, const int32_t *d_n_cars, double dt, double min_crash_distance,
                                            double *d_ego_out, double *d_cars_x_out, double *d_cars_v_out, double *d_cars_a_out,
                                            uint8_t *d_crashed, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_n_cars || !d_ego_out || !d_cars_x_out || !d_cars_v_out)
        return mpc_set_error(MPC_E_INVALID, "mpc_predict_step_without_ego: null pointer");
    MPC_CUDA_OK(launch_predict_step_without_ego(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, dt, min_crash_distance,
                                                d_ego_out, d_cars_x_out, d_cars_v_out, d_cars_a_out, d_crashed, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

extern "C" int mpc_state_vector(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                                const double *d_cars_a, const int32_t *d_n_cars, float *d_out, int out_stride, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_cars_a || !d_n_cars || !d_out || out_stride < 20) return mpc_set_error(MPC_E_INVALID, "mpc_state_vector: bad argument");
    MPC_CUDA_OK(launch_state_vector(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, d_out, out_stride, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

extern "C" int mpc_speed_from_jerk(mpc_handle *h, int B, const double *d_ego, const double *d_jerk, double *d_speed, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_jerk || !d_speed) return mpc_set_error(MPC_E_INVALID, "mpc_speed_from_jerk: null pointer");
    MPC_CUDA_OK(launch_speed_from_jerk(h->P, B, d_ego, d_jerk, d_speed, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

extern "C" int mpc_rollout_step(mpc_handle *h, int B, double *d_ego, double *d_cars_x, double *d_cars_v, double *d_cars_a,
                                const int32_t *d_n_cars, const double *d_jerk, double dt, double min_crash_distance, double stop_x,
                                int step, uint8_t *d_alive, double *d_selected_speed, double *d_roll_s, int roll_stride,
                                int32_t *d_roll_len, uint8_t *d_crash_predicted, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!d_ego || !d_cars_x || !d_cars_v || !d_cars_a || !d_n_cars || !d_jerk || !d_alive || !d_selected_speed || !d_roll_s ||
        !d_roll_len || !d_crash_predicted)
        return mpc_set_error(MPC_E_INVALID, "mpc_rollout_step: null pointer");
    if (step < 1 || step >= roll_stride) return mpc_set_error(MPC_E_INVALID, "mpc_rollout_step: step must be in [1, roll_stride)");
    MPC_CUDA_OK(launch_rollout_step(h->P, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, d_jerk, dt, min_crash_distance,
                                    stop_x, step, d_alive, d_selected_speed, d_roll_s, roll_stride, d_roll_len, d_crash_predicted,
                                    (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}

extern "C" int mpc_env_step(mpc_handle *h, const mpc_env_params *ep, int B, double *d_ego, double *d_cars_x, double *d_cars_v,
                            double *d_cars_a, int32_t *d_n_cars, double *d_prev_acc, double *d_delay, int32_t *d_ticks,
                            const double *d_jerk, const double *d_u_spawn, const double *d_fresh_gap_u, const double *d_fresh_first_u,
                            const double *d_fresh_speed_z, const double *d_fresh_delay_u, double *d_reward, uint8_t *d_flags,
                            double *d_projected_jerk, void *stream) {
    int rc = check_batch(h, B); if (rc) return rc;
    if (B == 0) return MPC_OK;
    if (!ep || !d_ego || !d_cars_x || !d_cars_v || !d_cars_a || !d_n_cars || !d_prev_acc || !d_delay || !d_ticks || !d_jerk || !d_reward ||
        !d_flags || !d_projected_jerk)
        return mpc_set_error(MPC_E_INVALID, "mpc_env_step: null pointer");
    if (ep->auto_reset && (!d_fresh_gap_u || !d_fresh_first_u || !d_fresh_delay_u))
        return mpc_set_error(MPC_E_INVALID, "mpc_env_step: auto_reset needs the fresh-episode random numbers");
    if (!(ep->tick > 0)) return mpc_set_error(MPC_E_INVALID, "mpc_env_step: tick must be positive");
    MPC_CUDA_OK(launch_env_step(h->P, *ep, B, h->nmax, d_ego, d_cars_x, d_cars_v, d_cars_a, d_n_cars, d_prev_acc, d_delay, d_ticks, d_jerk,
                                d_u_spawn, d_fresh_gap_u, d_fresh_first_u, d_fresh_speed_z, d_fresh_delay_u, d_reward, d_flags,
                                d_projected_jerk, (cudaStream_t)stream));
    h->kernels_launched = 1;
    return MPC_OK;
}
