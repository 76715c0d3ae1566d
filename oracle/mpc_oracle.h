/* CPU oracle for the S-T MPC hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
 *
 * A plain-C, fp64, single-problem restatement of the reference algorithm
 * (jlubars/RL-MPC-LaneMerging).  Every function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may link or call this.  The product path (libmpcb200.so) never does.
 *
 * Parity status: PINNED.  Checked bit-for-bit against the reference's own code run in the
 * build container (compiled st_cy.pyx + imported st.py/prediction.py/control.py/dqn.py) by
 * tests/golden/make_golden.py; the resulting vectors are committed under tests/golden/.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_CARS 64

/* Scalars the path reads from the reference's global Settings (config.py:7-155). */
typedef struct orc_params {
    double s_disc, t_disc, future_s, future_t;          /* S/T_DISCRETIZATION, FUTURE_S/T  */
    double start_uncertainty, uncertainty_per_second;
    double d_weight, v_weight, a_weight, j_weight;
    double desired_speed, max_speed;
    double a_min, a_max, j_min, j_max;                  /* MAX_NEGATIVE/POSITIVE_ACCELERATION, MINIMUM_NEGATIVE/MAXIMUM_POSITIVE_JERK */
    double min_allowed_distance, crash_min_s, car_length;
    double max_predicted_decel;                         /* MAX_PREDICTED_DECELERATION (-4) */
    double tick_length, sensor_radius;
    double combination_min_distance;
} orc_params;

/* prediction.HighwayState (prediction.py:9-20); others sorted front->back */
typedef struct orc_state {
    double ego_x, ego_y, ego_v, ego_a;
    int n;
    double x[ORC_MAX_CARS], v[ORC_MAX_CARS], a[ORC_MAX_CARS];
} orc_state;

void   orc_default_params(orc_params *p);
double orc_get_ego_s(double x, double y);                                   /* control.py:373-380 */
int    orc_arange_len(double start, double stop, double step);              /* numpy arange length */
void   orc_grid_dims(const orc_params *p, double start_s, int *num_t, int *num_s, double *delta_s_eff);

/* prediction.py:46-105 ; returns crashed flag, writes next state (may alias in) */
int  orc_predict_step_with_ego(const orc_params *p, const orc_state *in, double selected_speed,
                               double dt, double min_crash_distance, orc_state *out);
/* prediction.py:22-44 */
int  orc_predict_step_without_ego(const orc_params *p, const orc_state *in, double dt,
                                  double min_crash_distance, orc_state *out);

/* st.py:25-70.  obstacles[num_t*num_s] u8, distances[num_t*num_s] f64, s_values[num_s].
 * obs_s_out (optional, num_t*ORC_MAX_CARS, NaN-padded) receives the per-layer obstacle s of
 * every predicted car (whether or not it passed the range filters). */
void orc_build_grid(const orc_params *p, const orc_state *st, int num_t, int num_s,
                    uint8_t *obstacles, double *distances, double *s_values, double *obs_s_out);

typedef struct orc_solve_stats {
    long pops, pushes;          /* dijkstra */
    long nodes, edges;          /* layered */
    long coast_collisions;      /* layered: nodes sharing a coast cell with an earlier node of the layer */
    int  max_span, max_width;
    long wmult_hist[8];         /* layered: #window-start keys holding 1,2,..,8+ nodes (multimap sizing) */
} orc_solve_stats;

/* st_cy.pyx:315-399 (== st.py:361-452): Dijkstra with the reference's heap-tuple order.
 * idx_out[num_t] (int, -1 for unreached layers), s_seq_out[num_t] (0.0 for unreached),
 * returns reached_t (index of deepest layer on the returned path). cost_out = sum of edge costs. */
int orc_solve_dijkstra(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                       const double *distances, const double *s_values, double delta_t,
                       double v0, double a0, int *idx_out, double *s_seq_out, double *cost_out,
                       orc_solve_stats *stats);

/* Same answer by forward layered DP (SURVEY.md §7 "key design facts"); used as the fast CPU
 * baseline and to validate the DP==Dijkstra equivalence the CUDA kernels rely on. */
int orc_solve_layered(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                      const double *distances, const double *s_values, double delta_t,
                      double v0, double a0, int *idx_out, double *s_seq_out, double *cost_out,
                      orc_solve_stats *stats);

/* CPU model of the CUDA fast kernel's arithmetic (NOT the reference; see mpc_oracle.c). */
int orc_solve_fast_model(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                         const double *distances, const double *s_values, double delta_t,
                         double v0, double a0, int f32_labels, int *idx_out, double *s_seq_out, double *cost_out);

int orc_solve_fast_model_ex(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                            const double *distances, const double *s_values, double delta_t,
                            double v0, double a0, int f32_labels, uint64_t prune_fx,
                            int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts);

/* The 32-bit-key kernel's arithmetic (mpc_fast32.cu): labels in 2^-frac_bits fixed point; the min-combine orders candidates by
 * (label >> key_shift, larger v'), i.e. the low key_shift bits of a label are carried exactly but do not take part in comparisons. */
int orc_solve_fast_model_q(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                           const double *distances, const double *s_values, double delta_t,
                           double v0, double a0, int frac_bits, int key_shift, uint64_t prune_fx,
                           int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts);

/* Same model with a per-cell heuristic table hfx[num_t*num_s] (label units): a node of layer >= 2 is dropped when
 * label + hfx[cell] > prune_fx.  Checks the exactness of the reachability heuristic (oracle/bound_model.py).
 * fmin_out (optional, num_t entries): smallest label + h among the surviving nodes of each layer. */
int orc_solve_fast_model_h(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                           const double *distances, const double *s_values, double delta_t,
                           double v0, double a0, uint64_t prune_fx, const uint64_t *hfx, uint64_t *fmin_out,
                           int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts);

/* widest layer span (cells between the lowest and highest successor of a layer's surviving nodes) of the last
 * fast-model run on this thread of control: what the fast kernel's label ring must hold (not thread safe; tools only) */
int orc_last_max_span(void);

/* Sum of st.cost (st.py:140-144) along an index path with the solver's history convention. */
double orc_path_cost(const orc_params *p, int n, const int *idx, const double *s_values,
                     const double *distances, int num_s, double delta_t, double v0, double a0);

/* st.get_appropriate_base_st_path_and_obstacles (st.py:726-754) + st.test_guaranteed_crash (790-802).
 * use_layered: 0 = dijkstra, 1 = layered DP.  Returns reached_t. */
int orc_plan(const orc_params *p, const orc_state *st, int use_layered, int *idx_out,
             double *s_seq_out, double *cost_out, int *guaranteed_crash, double *min_path_distance,
             double *start_s_out, double *delta_s_out, int *num_s_out, orc_solve_stats *stats);

/* Batched + multi-threaded (pthread) plan over SoA arrays: CPU baseline. cars_* are [B,nmax]. */
void orc_plan_batch(const orc_params *p, int B, int nmax, const double *ego /*[B,4]*/,
                    const double *cars_x, const double *cars_v, const double *cars_a,
                    const int *n_cars, int use_layered, int nthreads, int num_t,
                    int *idx_out /*[B,num_t]*/, double *s_seq_out /*[B,num_t]*/, double *cost_out,
                    int *reached_out, int *crash_out, double *min_dist_out);

double orc_speed_from_jerk(const orc_params *p, double v, double a, double jerk);   /* control.py:160-171 */
void   orc_state_vector(const orc_params *p, const orc_state *st, double *out20);   /* dqn.py:389-446 */
double orc_path_mean_abs_jerk(const double *s, int n, double v0, double a0, double dt); /* st.py:274-288 */

#ifdef __cplusplus
}
#endif
#endif
