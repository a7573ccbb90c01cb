/* TEST INFRASTRUCTURE — not product code.
 *
 * C interface of the CPU oracle (cf_oracle.c): a plain-C restatement of the reference's
 * multi-agent predictive rollout (ghostplanner::cfplanner::CfManager / CfAgent,
 * /root/reference/src/bimanual_planning_ros/src/cf_manager.cpp, src/cf_agent.cpp).
 * The entry points mirror oracle/ref_harness.cpp one-to-one (prefix cforacle_ instead of
 * cfref_) so that tests can drive the real reference build, this restatement and the CUDA
 * product with the same script. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or load it.
 */
#ifndef CF_ORACLE_H
#define CF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

void *cforacle_create(void);
void cforacle_destroy(void *h);
void cforacle_init(void *h, const double *goal, double delta_t, int n_obs, const double *obs_pos,
                   const double *obs_vel, const double *obs_rad, int n_agents, const double *k_a,
                   const double *k_c, const double *k_r, const double *k_d, const double *k_manip,
                   int n_force, const double *k_r_force, double vel_max, double approach_dist,
                   double detect_shell_rad, unsigned long max_prediction_steps,
                   unsigned long prediction_freq_multiple, double agent_mass, double radius,
                   int keep_threads);
int cforacle_num_agents(void *h);
void cforacle_set_random_vecs(void *h, const double *vecs, int n_obs);
void cforacle_get_random_vecs(void *h, double *vecs, int n_obs);
void cforacle_set_initial_position(void *h, const double *p);
void cforacle_set_real_position(void *h, const double *p);
void cforacle_start_prediction(void *h);
void cforacle_stop_prediction(void *h);
double cforacle_rollout_threads(void *h);
double cforacle_rollout_pooled(void *h, int n_threads);
int cforacle_evaluate_agents(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                             const double *obs_rad, double k_goal_dist, double k_path_len,
                             double k_safe_dist, double k_workspace, const double *ws);
void cforacle_move_real_agent(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                              const double *obs_rad, double delta_t, int steps, int agent_id);
void cforacle_reset_agents(void *h, const double *pos, const double *vel, int n_obs,
                           const double *obs_pos, const double *obs_vel, const double *obs_rad);
void cforacle_get_next_position(void *h, double *p);
void cforacle_get_next_velocity(void *h, double *p);
void cforacle_get_ee_force(void *h, double *p);
double cforacle_get_dist_from_goal(void *h);
int cforacle_get_best_agent_type(void *h);
int cforacle_get_best_agent_id(void *h);
int cforacle_get_num_prediction_steps(void *h, int agent);
int cforacle_get_real_num_steps(void *h);
void cforacle_get_agent_summaries(void *h, int *steps, double *length, double *min_obs_dist,
                                  int *reached, double *pred_time_ns, int *agent_type);
void cforacle_get_predicted_paths(void *h, double *out, int stride);
void cforacle_get_agent_velocities(void *h, double *out);
int cforacle_get_planned_trajectory(void *h, double *out, int max_points);
void cforacle_get_obstacle_state(void *h, int n_obs, int *known, double *rot);
int cforacle_host_threads(void);
/* last evaluate's per-agent costs (the reference keeps them local to evaluateAgents) */
void cforacle_get_costs(void *h, double *costs);

#ifdef __cplusplus
}
#endif
#endif
