/* libmpcb200 -- B200 (sm_100a) native S-T MPC hot path, C ABI.
 *
 * Drop-in boundary for the one native component of jlubars/RL-MPC-LaneMerging
 * (the CPython extension `st_cy`, reference st_cy.pyx:315 `solve_s_t_path_fast`, called from
 * st.py:740-746) widened batch-first to the functions around it on the hot path
 * (SURVEY.md §8a/§8b).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every `d_*` pointer is a DEVICE pointer on the handle's device, every `h_*` pointer a HOST
 *    pointer (pinned memory gives asynchronous copies);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - all entry points return 0 on success or a negative MPC_E_* code; mpc_last_error() returns
 *    a thread-local message for the last failure;
 *  - an infeasible plan is NOT an error (reference st.py:762-767): it is reported as
 *    reached_t < num_t-1, idx = -1 and s_seq = 0.0 for the unreached layers;
 *  - traffic state layout (structure of arrays, fp64 like the reference's HighwayState,
 *    prediction.py:9-20):  ego[B][4] = (x, y, speed, acceleration); cars_x/v/a[B][nmax] sorted
 *    front->back (descending x), entries >= n_cars[b] ignored; n_cars[B] int32.
 *  - a handle is bound to one device and is not thread-safe; use one handle per stream.
 */
#ifndef MPCB200_H
#define MPCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_ABI_VERSION 1

enum {
    MPC_OK = 0,
    MPC_E_INVALID = -1,      /* bad argument */
    MPC_E_CUDA = -2,         /* CUDA runtime failure (message has the cudaError string) */
    MPC_E_CAPACITY = -3,     /* batch / nmax / grid larger than the handle was created for */
    MPC_E_NODEVICE = -4      /* no usable CUDA device: the library has NO CPU fallback */
};

/* arithmetic of the DP solve */
enum {
    MPC_MODE_FAST = 0,   /* integer-cell kinematics, 2^-18 fixed-point 48-bit labels (cost within 1e-7 rel) */
    MPC_MODE_EXACT = 1   /* fp64, reference operation order: index-identical to st_cy            */
};

/* Snapshot of the reference's global `Settings` scalars the path reads (config.py:30-37,94-110,
 * 143,146-153).  Taken at mpc_create / mpc_set_params; later Settings mutation needs a refresh. */
typedef struct mpc_params {
    double s_disc, t_disc, future_s, future_t;             /* S/T_DISCRETIZATION, FUTURE_S/T        */
    double start_uncertainty, uncertainty_per_second;
    double d_weight, v_weight, a_weight, j_weight;
    double desired_speed, max_speed;
    double a_min, a_max, j_min, j_max;                     /* MAX_NEGATIVE/POSITIVE_ACCELERATION, MINIMUM_NEGATIVE/MAXIMUM_POSITIVE_JERK */
    double min_allowed_distance, crash_min_s, car_length;
    double max_predicted_decel;                            /* MAX_PREDICTED_DECELERATION            */
    double tick_length, sensor_radius;
    double combination_min_distance;
} mpc_params;

typedef struct mpc_handle mpc_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */
int         mpc_abi_version(void);
const char *mpc_last_error(void);
void        mpc_default_params(mpc_params *p);           /* configs/st_*.json values               */
int         mpc_device_count(void);                      /* 0 when there is no GPU                 */
/* Scratch (layer descriptors, back-pointers, work counters) is sized here; nothing is allocated
 * on the hot path.  nmax <= 32. */
int mpc_create(const mpc_params *p, int device, int max_batch, int nmax, mpc_handle **out);
int mpc_set_params(mpc_handle *h, const mpc_params *p);
int mpc_destroy(mpc_handle *h);
/* num_t is fixed by the params; num_s depends on start_s through numpy.arange's length rule
 * (st.py:31), so only its maximum is a handle constant. */
int mpc_grid_dims(const mpc_handle *h, int *num_t, int *num_s_max);
/* Row stride (cells) of the dense grids mpc_build_grid writes: num_s_max rounded up to a multiple of 8 so that rows can be
 * written with vector stores. */
int mpc_grid_stride(const mpc_handle *h);
/* counters of the last call on this handle: [0] kernels launched, [1] problems that overflowed
 * the fast kernel's shared-memory window / bucket capacity and were re-solved by the exact kernel */
int mpc_last_counters(const mpc_handle *h, int64_t *out2);
/* The 32-bit-key DP kernel that makes the first (bounded) attempt of mpc_plan in MPC_MODE_FAST (csrc/mpc_fast32.cuh):
 * out6 = {in use (0/1), fraction bits of its fixed-point labels, its cost bound in label units, ring capacity in cells of its
 * first launch shape, problems of the last call it handed to the 64-bit kernel, problems of the last call its first launch
 * shape handed on (to its wide-ring shape or the 64-bit kernel)}.  The CPU model of the kernel (oracle/) takes the second
 * and third entry. */
int mpc_fast32_info(const mpc_handle *h, int64_t *out6);

/* Optional per-kernel device timing (used by bench.py for the roofline of the dominant kernel):
 * after mpc_set_timing(h,1), mpc_last_kernel_ms returns {traffic-predictor ms, DP-kernel ms,
 * fallback-DP ms} of the last mpc_plan / mpc_solve_dense call ({traffic-predictor ms, rasteriser ms, 0} after
 * mpc_build_grid), measured with CUDA events on the stream the kernels were launched on. */
int mpc_set_timing(mpc_handle *h, int enable);
int mpc_last_kernel_ms(mpc_handle *h, float *out3);

/* Self-test: rebuilds the layer descriptors for the given states and counts the cells whose sorted O(1)
 * obstacle/distance lookup (fast kernel) differs from the reference-order evaluation (must be 0). */
int mpc_selftest_search(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x, const double *d_cars_v,
                        const int32_t *d_n_cars, int64_t *mismatching_cells, void *stream);

/* ---- K1: traffic prediction + S-T rasterisation (st.find_s_t_obstacles_from_state, st.py:25-70,
 *      with prediction.py:22-105 and control.py:373-389) ------------------------------------- */
/* Dense grids, the layout st_cy consumes: d_obstacles u8[B][num_t][stride], d_distances f64 (dist_f32=0) or f32
 * (dist_f32=1) [B][num_t][stride] with stride = mpc_grid_stride(h); cells >= num_s[b] are
 * written as obstacle=1 / distance=0.  d_start_s f64[B], d_delta_s f64[B] (= s_values[1]-s_values[0]),
 * d_num_s i32[B] may each be NULL. */
int mpc_build_grid(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
                   const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars,
                   uint8_t *d_obstacles, void *d_distances, int dist_f32,
                   double *d_start_s, double *d_delta_s, int32_t *d_num_s, void *stream);

/* ---- K2: the drop-in for st_cy.solve_s_t_path_fast (st_cy.pyx:315-399) on caller-supplied dense
 *      grids.  s_values are implied by (start_s, delta_s, num_s) exactly as numpy.arange fills
 *      them.  Outputs: d_idx i32[B][num_t], d_s_seq f64[B][num_t], d_cost f64[B] (sum of edge
 *      costs along the returned path), d_reached_t i32[B]; any output may be NULL. ------------- */
int mpc_solve_dense(mpc_handle *h, int B, int num_t, int num_s_stride, const uint8_t *d_obstacles,
                    const void *d_distances, int dist_f32, const double *d_start_s,
                    const double *d_delta_s, const int32_t *d_num_s, const double *d_v0,
                    const double *d_a0, int mode, int32_t *d_idx, double *d_s_seq, double *d_cost,
                    int32_t *d_reached_t, void *stream);

/* ---- K3: fused gap-evaluation = st.get_appropriate_base_st_path_and_obstacles (st.py:726-754)
 *      + st.test_guaranteed_crash_from_state (st.py:790-802) without materialising the grid.
 *      Extra outputs: d_crash u8[B] (guaranteed-crash verdict), d_min_dist f64[B] (smallest
 *      distance-field value along the path, +inf when the plan is incomplete),
 *      d_start_s f64[B].  Any output may be NULL. ---------------------------------------------- */
int mpc_plan(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
             const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars, int mode,
             int32_t *d_idx, double *d_s_seq, double *d_cost, int32_t *d_reached_t,
             uint8_t *d_crash, double *d_min_dist, double *d_start_s, void *stream);

/* mpc_plan for the episodes with d_mask[b] != 0 only; the outputs of the other episodes are left untouched (d_start_s is
 * not available: pass NULL).  The masked episodes are listed and counted on the device, so the call involves no host round
 * trip: the combined controller (dqn.py:144-155) hands the vetoed episodes to st.do_st_control without synchronising. */
int mpc_plan_masked(mpc_handle *h, int B, const uint8_t *d_mask, const double *d_ego, const double *d_cars_x,
                    const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars, int mode,
                    int32_t *d_idx, double *d_s_seq, double *d_cost, int32_t *d_reached_t, uint8_t *d_crash,
                    double *d_min_dist, double *d_start_s, void *stream);

/* mpc_plan with a per-problem COST HINT for MPC_MODE_FAST (ignored by MPC_MODE_EXACT): an estimate of each plan's cost
 * -- the cost of a coarse probe plan (mpc_plan_probed), or of the previous tick's plan when the reference's
 * control loop re-plans every TICK_LENGTH (control.py:229-340 calls st.do_st_control once per tick).  Problem b
 * is first solved under the bound hint_scale * d_hint_cost[b] (when d_hint_reached_t == NULL or d_hint_reached_t[b] ==
 * hint_full_t; a hint that is <= 0, NaN or >= 1e9 is ignored).  The RESULT DOES NOT DEPEND ON THE HINT: a bounded
 * pass that reaches the horizon returns the unbounded answer, one that does not is repeated under the standard
 * bound; a hint only changes how many nodes the solve expands.  The hint arrays must not alias the
 * outputs. */
int mpc_plan_hinted(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
                    const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars, int mode,
                    const double *d_hint_cost, const int32_t *d_hint_reached_t, int hint_full_t,
                    double hint_scale, int32_t *d_idx, double *d_s_seq, double *d_cost,
                    int32_t *d_reached_t, uint8_t *d_crash, double *d_min_dist, double *d_start_s,
                    void *stream);

/* Probe + plan in one call (MPC_MODE_FAST): `probe` is a second handle on the same device, created from the same
 * Settings with a coarser S_DISCRETIZATION / T_DISCRETIZATION (e.g. 20x / 3x: an 18 x 451 grid instead of 51 x 9001).
 * The states are planned on the probe grid first; margin * (num_t-1)/(probe num_t-1) * probe cost is the hint of
 * the real solve.  Outputs as mpc_plan (identical to mpc_plan's, see above). */
int mpc_plan_probed(mpc_handle *h, mpc_handle *probe, double margin, int B, const double *d_ego,
                    const double *d_cars_x, const double *d_cars_v, const double *d_cars_a,
                    const int32_t *d_n_cars, int32_t *d_idx, double *d_s_seq, double *d_cost,
                    int32_t *d_reached_t, uint8_t *d_crash, double *d_min_dist, double *d_start_s,
                    void *stream);

/* Same call with HOST buffers: copies the state in, plans, copies the results out and
 * synchronises the stream (the end-to-end path a CPU caller such as the reference's
 * control.run_episode loop uses). */
int mpc_plan_host(mpc_handle *h, int B, const double *h_ego, const double *h_cars_x,
                  const double *h_cars_v, const double *h_cars_a, const int32_t *h_n_cars, int mode,
                  int32_t *h_idx, double *h_s_seq, double *h_cost, int32_t *h_reached_t,
                  uint8_t *h_crash, double *h_min_dist, double *h_start_s, void *stream);

/* mpc_plan_probed with HOST buffers (the end-to-end form of the probed solve). */
int mpc_plan_host_probed(mpc_handle *h, mpc_handle *probe, double margin, int B, const double *h_ego,
                         const double *h_cars_x, const double *h_cars_v, const double *h_cars_a,
                         const int32_t *h_n_cars, int32_t *h_idx, double *h_s_seq, double *h_cost,
                         int32_t *h_reached_t, uint8_t *h_crash, double *h_min_dist, double *h_start_s,
                         void *stream);

/* ---- finer_fit (st.py:584-723): re-sample each plan from T_DISCRETIZATION to TICK_LENGTH by linear interpolation and
 * project it onto the speed / acceleration / jerk limits (the QP the reference gives to cvxopt; solved here by a
 * primal-dual interior-point method, one thread per episode).  d_s_seq f64[B][num_t] and d_reached_t i32[B] are mpc_plan's
 * outputs; d_fine f64[B][fine_stride] receives the smoothed positions, d_n_fine i32[B] their count (1 = the plan had a
 * single point, st.py:587-588), d_speed f64[B] (optional) the speed command (fine[1]-fine[0])/TICK_LENGTH of
 * st.py:780-781 (the current ego speed when the plan has a single point, st.py:775-777), d_iterations i32[B] (optional). */
int mpc_finer_fit(mpc_handle *h, int B, const double *d_s_seq, const int32_t *d_reached_t, const double *d_ego,
                  double *d_fine, int fine_stride, int32_t *d_n_fine, double *d_speed, int32_t *d_iterations, void *stream);
/* the same for the episodes with d_mask[b] != 0 only (rows of the others untouched, no host round trip) */
int mpc_finer_fit_masked(mpc_handle *h, int B, const uint8_t *d_mask, const double *d_s_seq, const int32_t *d_reached_t,
                         const double *d_ego, double *d_fine, int fine_stride, int32_t *d_n_fine, double *d_speed,
                         int32_t *d_iterations, void *stream);
int mpc_finer_fit_max_points(void);

/* ---- K4: rollout tick pieces ------------------------------------------------------------------
 * HighwayState.predict_step_with_ego (prediction.py:46-105), batched, in place allowed
 * (out pointers may alias in pointers).  d_selected_speed f64[B]; d_crashed u8[B]. */
int mpc_predict_step_with_ego(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
                              const double *d_cars_v, const double *d_cars_a,
                              const int32_t *d_n_cars, const double *d_selected_speed, double dt,
                              double min_crash_distance, double *d_ego_out, double *d_cars_x_out,
                              double *d_cars_v_out, double *d_cars_a_out, uint8_t *d_crashed,
                              void *stream);
/* One tick of the SUMO-free world with the car-following model of the reference's traffic (merge_impossible.rou.xml:3: Krauss,
 * accel / decel / tau / minGap; max_speed = OTHER_CAR_SPEED, sumo.py:60), IN PLACE: every car follows its leader (the next car
 * ahead, or the ego once it is on the junction / highway), the ego moves at d_selected_speed as in predict_step_with_ego, whose
 * crash test is applied.  Alternative to mpc_predict_step_with_ego as the dynamics of merge_gym.MergeEnv (Settings.WORLD_MODEL). */
int mpc_krauss_step(mpc_handle *h, int B, double *d_ego, double *d_cars_x, double *d_cars_v, double *d_cars_a,
                    const int32_t *d_n_cars, const double *d_selected_speed, double dt, double min_crash_distance,
                    double accel, double decel, double tau, double min_gap, double max_speed, uint8_t *d_crashed, void *stream);
/* HighwayState.predict_step_without_ego (prediction.py:22-44): the traffic-only step the S-T grid builder applies per
 * layer (st.py:42-43) -- the ego is replaced by a pseudo-ego (stays put before the merge point, ignored when it leads all
 * cars, otherwise placed CAR_LENGTH + 5 m behind the car ahead of it at that car's speed) and the state advances with
 * predict_step_with_ego.  Same buffers and aliasing rules as mpc_predict_step_with_ego. */
int mpc_predict_step_without_ego(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
                                 const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars,
                                 double dt, double min_crash_distance, double *d_ego_out,
                                 double *d_cars_x_out, double *d_cars_v_out, double *d_cars_a_out,
                                 uint8_t *d_crashed, void *stream);
/* dqn.get_state_vector_from_base_state (dqn.py:389-446) -> f32 [B][out_stride] (first 20 columns
 * written); control.get_ego_speed_from_jerk (control.py:160-171) -> f64[B]. */
int mpc_state_vector(mpc_handle *h, int B, const double *d_ego, const double *d_cars_x,
                     const double *d_cars_v, const double *d_cars_a, const int32_t *d_n_cars,
                     float *d_out, int out_stride, void *stream);
int mpc_speed_from_jerk(mpc_handle *h, int B, const double *d_ego, const double *d_jerk,
                        double *d_speed, void *stream);
/* One step of RLAgent.do_combined_control's policy rollout (dqn.py:129-141), masked and IN PLACE: for every episode
 * with d_alive[b] != 0 the proposed jerk becomes a speed (control.py:160-171), the state advances with
 * predict_step_with_ego, d_selected_speed[b] is that speed, d_roll_s[b][step] the new ego arclength
 * (rollout_s_history), d_roll_len[b] += 1, d_crash_predicted[b] |= crashed, and d_alive[b] is cleared when the step
 * crashed or passed stop_x (Settings.STOP_X).  Finished episodes keep their state; d_roll_s[b][step] repeats
 * d_roll_s[b][step-1].  1 <= step < roll_stride. */
int mpc_rollout_step(mpc_handle *h, int B, double *d_ego, double *d_cars_x, double *d_cars_v, double *d_cars_a,
                     const int32_t *d_n_cars, const double *d_jerk, double dt, double min_crash_distance,
                     double stop_x, int step, uint8_t *d_alive, double *d_selected_speed, double *d_roll_s,
                     int roll_stride, int32_t *d_roll_len, uint8_t *d_crash_predicted, void *stream);

/* ---- one tick of the batched lane-merging environment, fused (the per-tick body of the reference's
 * ContinuousJerkEnv.step, merge_gym.py:83-162, on the SUMO-free world of rl_mpc_lanemerging_b200/merge_gym.py): invalid-action
 * clipping (merge_gym.py:83-96), the world step (predict_step_with_ego as dynamics), recycling of the front car beyond
 * sensor range, entry of a new car every BASE_TRAFFIC_INTERVAL + U[0,1) s (control.py:215-226), termination (arrived /
 * collision / timeout), the Slotted-Jerk reward (dqn.py:557-563) and, with auto_reset, fresh initial conditions for the
 * finished episodes.  One warp per episode, everything in place.  Random numbers are INPUTS (the caller draws them with its
 * own generator): d_u_spawn f64[B] (U[0,1), NULL = 0), and for auto_reset d_fresh_gap_u f64[B][nmax], d_fresh_first_u f64[B],
 * d_fresh_speed_z f64[B] (N(0,1); NULL = no start-speed randomisation), d_fresh_delay_u f64[B].
 * Outputs: d_reward f64[B], d_flags u8[4][B] = (done, crashed, arrived, timeout), d_projected_jerk f64[B]. */
typedef struct mpc_env_params {
    double tick, a_min, a_max, max_speed, min_crash_distance, sensor_radius, spawn_x, other_speed, interval, arrival_x;
    double ego_start_x, ego_start_y, start_speed, start_speed_var, min_start_speed, max_start_speed;
    double time_reward_step;        /* TIME_REWARD * TICK_LENGTH */
    double jerk_weight, crash_reward, success_reward;
    double invalid_action_step;     /* INVALID_ACTION_PENALTY * TICK_LENGTH, added to the reward of a tick whose action was clipped (merge_gym.py:86-92) */
    double krauss_accel, krauss_decel, krauss_tau, krauss_min_gap;   /* world == 1: the traffic's Krauss model (mpc_krauss_step) */
    int32_t max_ticks, auto_reset;
    int32_t world, pad_;            /* 0 = the reference's predictor as dynamics (mpc_predict_step_with_ego), 1 = Krauss traffic */
} mpc_env_params;
int mpc_env_step(mpc_handle *h, const mpc_env_params *ep, int B, double *d_ego, double *d_cars_x, double *d_cars_v,
                 double *d_cars_a, int32_t *d_n_cars, double *d_prev_acc, double *d_delay, int32_t *d_ticks,
                 const double *d_jerk, const double *d_u_spawn, const double *d_fresh_gap_u,
                 const double *d_fresh_first_u, const double *d_fresh_speed_z, const double *d_fresh_delay_u,
                 double *d_reward, uint8_t *d_flags, double *d_projected_jerk, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPCB200_H */
