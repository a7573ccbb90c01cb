/* TEST INFRASTRUCTURE — not product code.
 *
 * CPU oracle: a plain-C restatement of the reference's multi-agent predictive rollout.
 * All citations are file:line under /root/reference/src/bimanual_planning_ros/.
 *
 *   CfAgent family    include/bimanual_planning_ros/cf_agent.h, src/cf_agent.cpp
 *   CfManager         include/bimanual_planning_ros/cf_manager.h, src/cf_manager.cpp
 *
 * Third-party arithmetic on the path: Eigen3 fixed-size 3-vectors only (system package, not
 * vendored, version unpinned by the reference; Eigen 3.3.7 semantics assumed — see
 * shim/eigen3/Eigen/Dense for the reduction order, normalisation and division rules restated
 * here in v3_*).
 *
 * PARITY PIN: the reference ships no tests, golden vectors or known answers for this path
 * (SURVEY.md §4), so this file is pinned against the reference ITSELF: oracle/_ref/libcfref.so
 * is the reference's unmodified cf_agent.cpp/cf_manager.cpp built in this container, and
 * tests/test_oracle_vs_ref.py requires bit-identical outputs from the two on the anchor task
 * and on randomised cases; tests/golden/ holds vectors generated from libcfref.so
 * (tests/golden/make_golden.py) that this file must reproduce bit-for-bit wherever the
 * reference build is not available.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: default x86-64 codegen of the reference).
 */
#include "cf_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- Eigen 3-vector semantics ------------------------------------------------------------ */
typedef struct {
  double x, y, z;
} v3;

static v3 v3_make(double x, double y, double z) {
  v3 r = {x, y, z};
  return r;
}
static v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 v3_scale(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static v3 v3_div(v3 a, double s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static double v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static double v3_norm(v3 a) { return sqrt(v3_dot(a, a)); }
static v3 v3_normalized(v3 a) {
  double z = v3_dot(a, a);
  if (z > 0.0) return v3_div(a, sqrt(z));
  return a;
}
static v3 v3_cross(v3 a, v3 b) {
  return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* ---- agent / manager state ----------------------------------------------------------------- */
/* CfAgent::Type, cf_agent.h:59-68 */
enum { REAL_AGENT = 0, GOAL_H, OBSTACLE_H, GOAL_OBSTACLE_H, VEL_H, RANDOM_A, HAD_H, UNDEFINED_T };

typedef struct {
  v3 *pos, *vel;
  double *rad;
  int n;
} obstacles_t;

typedef struct { /* cf_agent.h:34-56 (+ random_vecs_, :326) */
  int id, type;
  v3 *path;
  int n_path, cap_path;
  v3 vel, init_pos, goal, force;
  double shell, mass, rad, vel_max, min_obs_dist, approach_dist;
  obstacles_t obs; /* private copy (obstacles_) */
  int n_state;     /* size of known / rot / random */
  unsigned char *known;
  v3 *rot;
  v3 *random_vecs;
  double prediction_time;
  int reached_goal;
} agent_t;

typedef struct { /* what RealCfAgent needs from *best_agent_ (cf_agent.cpp:368-387) */
  int present, id, type, n_state;
  v3 *random_vecs;
} best_t;

typedef struct { /* cf_manager.h:19-34 */
  agent_t real;
  best_t best;
  agent_t *ee;
  int n_ee;
  double *k_a, *k_c, *k_r, *k_d, *k_manip;
  v3 init_pos, goal;
  double approach_dist;
  double pred_dt;
  size_t max_steps;
  double *costs;
} mgr_t;

static void obstacles_free(obstacles_t *o) {
  free(o->pos), free(o->vel), free(o->rad);
  memset(o, 0, sizeof *o);
}

static obstacles_t obstacles_from(int n, const double *pos, const double *vel, const double *rad) {
  obstacles_t o;
  o.n = n;
  o.pos = malloc(sizeof(v3) * (n ? n : 1));
  o.vel = malloc(sizeof(v3) * (n ? n : 1));
  o.rad = malloc(sizeof(double) * (n ? n : 1));
  for (int i = 0; i < n; ++i) {
    o.pos[i] = v3_make(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    o.vel[i] = v3_make(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    o.rad[i] = rad[i];
  }
  return o;
}

static void agent_free(agent_t *a) {
  free(a->path), free(a->known), free(a->rot), free(a->random_vecs);
  obstacles_free(&a->obs);
  memset(a, 0, sizeof *a);
}

static void path_push(agent_t *a, v3 p) {
  if (a->n_path == a->cap_path) {
    a->cap_path = a->cap_path ? 2 * a->cap_path : 64;
    a->path = realloc(a->path, sizeof(v3) * a->cap_path);
  }
  a->path[a->n_path++] = p;
}

static v3 latest(const agent_t *a) { return a->path[a->n_path - 1]; }

/* CfAgent constructor, cf_agent.h:69-97 */
static void agent_construct(agent_t *a, int id, int type, v3 agent_pos, v3 goal, double shell, double mass,
                            double radius, double vel_max, double approach, int num_obstacles,
                            const obstacles_t *obs) {
  memset(a, 0, sizeof *a);
  a->id = id, a->type = type;
  path_push(a, agent_pos);
  a->vel = v3_make(0.01, 0.0, 0.0);
  a->init_pos = v3_make(0.0, 0.0, 0.0);
  a->goal = goal;
  a->shell = shell, a->min_obs_dist = shell;
  a->mass = mass, a->rad = radius, a->vel_max = vel_max, a->approach_dist = approach;
  a->n_state = num_obstacles;
  a->known = calloc(num_obstacles ? num_obstacles : 1, 1);
  a->rot = malloc(sizeof(v3) * (num_obstacles ? num_obstacles : 1));
  for (int i = 0; i < num_obstacles; ++i) a->rot[i] = v3_make(0.0, 0.0, 1.0); /* :92-96 */
  if (obs) {
    a->obs.n = obs->n;
    a->obs.pos = malloc(sizeof(v3) * (obs->n ? obs->n : 1));
    a->obs.vel = malloc(sizeof(v3) * (obs->n ? obs->n : 1));
    a->obs.rad = malloc(sizeof(double) * (obs->n ? obs->n : 1));
    memcpy(a->obs.pos, obs->pos, sizeof(v3) * obs->n);
    memcpy(a->obs.vel, obs->vel, sizeof(v3) * obs->n);
    memcpy(a->obs.rad, obs->rad, sizeof(double) * obs->n);
  }
  if (type == RANDOM_A) { /* cf_agent.h:338-342; values are overwritten by set_random_vecs */
    a->random_vecs = malloc(sizeof(v3) * (num_obstacles ? num_obstacles : 1));
    for (int i = 0; i < num_obstacles; ++i) a->random_vecs[i] = v3_normalized(v3_make(1.0, 1.0, 1.0));
  }
}

static double dist_from_goal(const agent_t *a) { return v3_norm(v3_sub(a->goal, latest(a))); } /* cf_agent.h:119-121 */

/* ---- per-type heuristics --------------------------------------------------------------------- */
/* nearest other obstacle among 0..O-2, cf_agent.cpp:434-446 / :480-492 */
static int closest_other_obstacle(const obstacles_t *obs, int id) {
  double min_dist = 100.0;
  int closest = 0;
  for (int i = 0; i < obs->n - 1; ++i) {
    if (i != id) {
      double d = v3_norm(v3_sub(obs->pos[id], obs->pos[i]));
      if (min_dist > d) {
        min_dist = d;
        closest = i;
      }
    }
  }
  return closest;
}

/* calculateRotationVector of the agent `type`; random_vecs is that agent's table (RANDOM only). */
static v3 rotation_vector(int type, const v3 *random_vecs, v3 agent_pos, v3 goal_pos, const obstacles_t *obs,
                          int id) {
  switch (type) {
    case GOAL_H: /* cf_agent.cpp:408-412 */
    case VEL_H:  /* :539-543 */
      return v3_make(0.0, 0.0, 1.0);
    case OBSTACLE_H: { /* :428-461 */
      if (obs->n < 2) return v3_make(0.0, 0.0, 1.0);
      int c = closest_other_obstacle(obs, id);
      v3 obstacle_vec = v3_sub(obs->pos[c], obs->pos[id]);
      v3 to_obs = v3_normalized(v3_sub(obs->pos[id], agent_pos));
      v3 current = v3_sub(v3_scale(to_obs, v3_dot(obstacle_vec, to_obs)), obstacle_vec);
      return v3_normalized(v3_cross(current, to_obs));
    }
    case GOAL_OBSTACLE_H: { /* :477-518 */
      int c = closest_other_obstacle(obs, id);
      v3 obstacle_vec = v3_sub(obs->pos[c], obs->pos[id]);
      v3 to_obs = v3_normalized(v3_sub(obs->pos[id], agent_pos));
      v3 obst_current = v3_sub(v3_scale(to_obs, v3_dot(obstacle_vec, to_obs)), obstacle_vec);
      v3 goal_vec = v3_sub(goal_pos, agent_pos);
      v3 goal_current = v3_sub(goal_vec, v3_scale(to_obs, v3_dot(to_obs, goal_vec)));
      v3 current = v3_add(v3_normalized(goal_current), v3_normalized(obst_current));
      if (v3_norm(current) < 1e-10) current = v3_make(0.0, 0.0, 1.0);
      current = v3_normalized(current);
      return v3_normalized(v3_cross(current, to_obs));
    }
    case RANDOM_A: { /* :559-566 — not normalised */
      v3 goal_vec = v3_normalized(v3_sub(goal_pos, agent_pos));
      return v3_cross(goal_vec, random_vecs[id]);
    }
    case HAD_H: { /* :599-611 — NaN when d is parallel to goal_vec */
      v3 obs_pos = obs->pos[id];
      v3 goal_vec = v3_sub(goal_pos, agent_pos);
      v3 rob_obs = v3_sub(obs_pos, agent_pos);
      double gn = v3_norm(goal_vec);
      double s = v3_dot(rob_obs, goal_vec) / (gn * gn);
      v3 d = v3_sub(v3_add(agent_pos, v3_scale(goal_vec, s)), obs_pos);
      v3 c = v3_cross(d, goal_vec);
      return v3_div(c, v3_norm(c));
    }
    default: /* base-class virtual has no return statement (cf_agent.h:155-162): undefined */
      return v3_make(0.0, 0.0, 1.0);
  }
}

/* currentVector of the agent `type`; agent_vel is the RELATIVE velocity the caller passes (:100) */
static v3 current_vector(int type, v3 agent_pos, v3 agent_vel, v3 goal_pos, const obstacles_t *obs, int id,
                         const v3 *rot) {
  switch (type) {
    case GOAL_H: { /* :389-406 */
      v3 goal_vec = v3_sub(goal_pos, agent_pos);
      v3 to_obs = v3_normalized(v3_sub(obs->pos[id], agent_pos));
      v3 current = v3_sub(goal_vec, v3_scale(to_obs, v3_dot(to_obs, goal_vec)));
      if (v3_norm(current) < 1e-10) current = v3_make(0.0, 0.0, 1.0);
      return v3_normalized(current);
    }
    case VEL_H: { /* :520-537 */
      v3 nv = v3_normalized(agent_vel);
      v3 to_obs = v3_normalized(v3_sub(obs->pos[id], agent_pos));
      v3 current = v3_sub(nv, v3_scale(to_obs, v3_dot(nv, to_obs)));
      if (v3_norm(current) < 1e-10) current = v3_make(0.0, 0.0, 1.0);
      return v3_normalized(current);
    }
    case OBSTACLE_H:      /* :414-426 */
    case GOAL_OBSTACLE_H: /* :463-475 */
    case RANDOM_A:        /* :545-557 */
    case HAD_H: {         /* :585-597 */
      v3 to_obs = v3_normalized(v3_sub(obs->pos[id], agent_pos));
      return v3_normalized(v3_cross(to_obs, rot[id]));
    }
    default:
      return v3_make(0.0, 0.0, 0.0);
  }
}

/* ---- forces ------------------------------------------------------------------------------------ */
/* CfAgent::circForce (cf_agent.cpp:72-108) and RealCfAgent::circForce (:110-144). For the real
 * agent the heuristics are those of `heur_type` / `heur_random` (the best agent) and
 * min_obs_dist_ is left alone (:121-124). */
static void circ_force(agent_t *a, const obstacles_t *obs, double k_circ, int heur_type, const v3 *heur_random,
                       int is_real) {
  v3 p = latest(a);
  v3 goal_vec = v3_sub(a->goal, p);
  for (int i = 0; i < obs->n - 1; ++i) {
    v3 rov = v3_sub(obs->pos[i], p);
    v3 rel_vel = v3_sub(a->vel, obs->vel[i]);
    if (v3_dot(v3_normalized(rov), v3_normalized(goal_vec)) < -0.01 && v3_dot(rov, rel_vel) < -0.01) continue;
    /* |o - p| (:83) and |p - o| (:121) square the same magnitudes: identical */
    double dist_obs = v3_norm(rov) - (a->rad + obs->rad[i]);
    dist_obs = dist_obs < 1e-5 ? 1e-5 : dist_obs; /* std::max(dist_obs, 1e-5); NaN stays NaN */ /* std::max(dist_obs, 1e-5) */
    if (!is_real && dist_obs < a->min_obs_dist) a->min_obs_dist = dist_obs;
    v3 curr_force = v3_make(0.0, 0.0, 0.0);
    if (dist_obs < a->shell) {
      if (!a->known[i]) {
        a->rot[i] = rotation_vector(heur_type, heur_random, p, a->goal, obs, i);
        a->known[i] = 1;
      }
      double vel_norm = v3_norm(rel_vel);
      if (vel_norm != 0) {
        v3 nv = v3_div(rel_vel, vel_norm);
        v3 current = current_vector(heur_type, p, rel_vel, a->goal, obs, i, a->rot);
        curr_force = v3_scale(v3_cross(nv, v3_cross(current, nv)), k_circ / (dist_obs * dist_obs));
      }
    }
    a->force = v3_add(a->force, curr_force);
  }
}

/* cf_agent.cpp:159-181 — last obstacle only */
static void repel_force(agent_t *a, const obstacles_t *obs, double k_repel) {
  v3 p = latest(a);
  v3 back = obs->pos[obs->n - 1];
  v3 dist_vec = v3_sub(p, back); /* -(back - p): negation is exact, same magnitudes */
  double dist_obs = v3_norm(dist_vec) - (a->rad + obs->rad[obs->n - 1]);
  dist_obs = dist_obs < 1e-5 ? 1e-5 : dist_obs; /* std::max(dist_obs, 1e-5); NaN stays NaN */
  v3 repel = v3_make(0.0, 0.0, 0.0);
  if (dist_obs < a->shell) {
    v3 u = v3_normalized(v3_sub(p, back));
    double s1 = 1.0 / dist_obs - 1.0 / a->shell;
    double s2 = dist_obs * dist_obs;
    repel = v3_make(k_repel * u.x * s1 / s2, k_repel * u.y * s1 / s2, k_repel * u.z * s1 / s2);
  }
  /* total_repel_force = 0 + repel; force_ += total (:179-180) */
  v3 total = v3_add(v3_make(0.0, 0.0, 0.0), repel);
  a->force = v3_add(a->force, total);
}

/* cf_agent.cpp:183-193 */
static void attractor_force(agent_t *a, double k_attr, double k_damp, double k_goal_scale) {
  if (k_attr == 0.0) return;
  v3 goal_vec = v3_sub(a->goal, latest(a));
  v3 vel_des = v3_scale(goal_vec, k_attr / k_damp);
  double lim = a->vel_max / v3_norm(vel_des);
  double scale_lim = lim < 1.0 ? lim : 1.0; /* std::min(1.0, lim) = (lim < 1.0) ? lim : 1.0; NaN -> 1.0 */
  vel_des = v3_scale(vel_des, scale_lim);
  double k = k_goal_scale * k_damp;
  a->force = v3_add(a->force, v3_scale(v3_sub(vel_des, a->vel), k));
}

/* cf_agent.cpp:195-227 */
static double attractor_force_scaling(const agent_t *a, const obstacles_t *obs) {
  int id_closest = 0, no_close = 1;
  double closest = a->shell;
  v3 p = latest(a);
  for (int i = 0; i < obs->n - 1; ++i) {
    double d = v3_norm(v3_sub(p, obs->pos[i])) - (a->rad + obs->rad[i]);
    d = d < 1e-5 ? 1e-5 : d;
    if (d < closest) {
      no_close = 0;
      closest = d;
      id_closest = i;
    }
  }
  if (no_close) return 1;
  v3 goal_vec = v3_sub(a->goal, p);
  if (v3_dot(goal_vec, a->vel) <= 0.0 && v3_norm(a->vel) < a->vel_max - 0.1 * a->vel_max &&
      v3_norm(goal_vec) > 0.15)
    return 0.0;
  double w1 = 1 - exp(-sqrt(closest) / a->shell);
  v3 rov = v3_sub(obs->pos[id_closest], p);
  double w2 = 1 - (v3_dot(goal_vec, rov) / (v3_norm(goal_vec) * v3_norm(rov)));
  w2 = w2 * w2;
  return w1 * w2;
}

/* cf_agent.cpp:253-268 */
static void update_position_and_velocity(agent_t *a, double dt) {
  v3 acc = v3_div(a->force, a->mass);
  double acc_norm = v3_norm(acc);
  if (acc_norm > 13.0) acc = v3_scale(acc, 13.0 / acc_norm);
  v3 p = latest(a);
  v3 new_pos = v3_make((p.x + 0.5 * acc.x * dt * dt) + a->vel.x * dt, (p.y + 0.5 * acc.y * dt * dt) + a->vel.y * dt,
                       (p.z + 0.5 * acc.z * dt * dt) + a->vel.z * dt);
  a->vel = v3_add(a->vel, v3_scale(acc, dt));
  double vel_norm = v3_norm(a->vel);
  if (vel_norm > a->vel_max) a->vel = v3_scale(a->vel, a->vel_max / vel_norm);
  path_push(a, new_pos);
}

/* cf_agent.cpp:270-276 */
static void predict_obstacles(agent_t *a, double dt) {
  for (int i = 0; i < a->obs.n; ++i) a->obs.pos[i] = v3_add(a->obs.pos[i], v3_scale(a->obs.vel[i], dt));
}

/* the gate shared by cfPlanner / cfPrediction (cf_agent.cpp:287-289, :315-317, :352-354) */
static int field_gate_open(const agent_t *a) {
  return !(dist_from_goal(a) < a->approach_dist ||
           (v3_norm(a->vel) < 0.5 * a->vel_max && v3_norm(v3_sub(latest(a), a->init_pos)) < 0.2));
}

/* inner loop of CfAgent::cfPrediction, cf_agent.cpp:310-336, run to termination */
static void rollout(agent_t *a, double k_attr, double k_circ, double k_repel, double k_damp, double dt,
                    size_t max_steps) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int ran = 0;
  while (dist_from_goal(a) > 0.1 && (size_t)a->n_path < max_steps) {
    ran = 1;
    a->force = v3_make(0.0, 0.0, 0.0);
    double k_goal_scale = 1.0;
    if (field_gate_open(a)) {
      circ_force(a, &a->obs, k_circ, a->type, a->random_vecs, 0);
      if (v3_norm(a->force) > 1e-5) k_goal_scale = attractor_force_scaling(a, &a->obs);
    }
    repel_force(a, &a->obs, k_repel);
    attractor_force(a, k_attr, k_damp, k_goal_scale);
    update_position_and_velocity(a, dt);
    predict_obstacles(a, dt);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (ran) { /* `if (running_)`, :330-337 */
    a->prediction_time = (t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec);
    a->reached_goal = dist_from_goal(a) < 0.100001;
  }
}

/* ---- manager ------------------------------------------------------------------------------------- */
void *cforacle_create(void) { return calloc(1, sizeof(mgr_t)); }

static void mgr_clear_agents(mgr_t *m) {
  for (int i = 0; i < m->n_ee; ++i) agent_free(&m->ee[i]);
  free(m->ee);
  m->ee = NULL, m->n_ee = 0;
  free(m->k_a), free(m->k_c), free(m->k_r), free(m->k_d), free(m->k_manip), free(m->costs);
  m->k_a = m->k_c = m->k_r = m->k_d = m->k_manip = m->costs = NULL;
}

void cforacle_destroy(void *h) {
  mgr_t *m = h;
  mgr_clear_agents(m);
  agent_free(&m->real);
  free(m->best.random_vecs);
  free(m);
}

static double *dup_d(const double *s, int n) {
  double *d = malloc(sizeof(double) * (n ? n : 1));
  memcpy(d, s, sizeof(double) * n);
  return d;
}

/* CfManager::init, cf_manager.cpp:41-124. best_agent_ is NOT reset (quirk 6). */
void cforacle_init(void *h, const double *goal, double delta_t, int n_obs, const double *obs_pos,
                   const double *obs_vel, const double *obs_rad, int n_agents, const double *k_a,
                   const double *k_c, const double *k_r, const double *k_d, const double *k_manip,
                   int n_force, const double *k_r_force, double vel_max, double approach_dist,
                   double detect_shell_rad, unsigned long max_prediction_steps,
                   unsigned long prediction_freq_multiple, double agent_mass, double radius,
                   int keep_threads) {
  (void)n_force, (void)k_r_force, (void)keep_threads; /* force_agents_ have no caller on this path */
  mgr_t *m = h;
  mgr_clear_agents(m);
  m->goal = v3_make(goal[0], goal[1], goal[2]);
  m->k_a = dup_d(k_a, n_agents), m->k_c = dup_d(k_c, n_agents), m->k_r = dup_d(k_r, n_agents);
  m->k_d = dup_d(k_d, n_agents), m->k_manip = dup_d(k_manip, n_agents);
  m->approach_dist = approach_dist;
  m->pred_dt = prediction_freq_multiple * delta_t; /* :122 */
  m->max_steps = max_prediction_steps;
  obstacles_t obs = obstacles_from(n_obs, obs_pos, obs_vel, obs_rad);

  agent_free(&m->real); /* :66-68: a fresh RealCfAgent with no obstacle copy */
  agent_construct(&m->real, 0, REAL_AGENT, m->init_pos, m->goal, detect_shell_rad, agent_mass, radius, vel_max,
                  approach_dist, n_obs, NULL);

  /* :70-104 — at least the HAD agent is always created; type order HAD, GOAL, OBSTACLE,
   * GOAL_OBSTACLE, VEL, then RANDOM; ids = index + 1 */
  static const int order[5] = {HAD_H, GOAL_H, OBSTACLE_H, GOAL_OBSTACLE_H, VEL_H};
  int n = n_agents < 1 ? 1 : n_agents;
  m->ee = calloc(n, sizeof(agent_t));
  m->n_ee = n;
  for (int i = 0; i < n; ++i) {
    int type = i < 5 ? order[i] : RANDOM_A;
    agent_construct(&m->ee[i], i + 1, type, m->init_pos, m->goal, detect_shell_rad, agent_mass, radius, vel_max,
                    approach_dist, n_obs, &obs);
  }
  m->costs = calloc(n, sizeof(double));
  obstacles_free(&obs);
}

int cforacle_num_agents(void *h) { return ((mgr_t *)h)->n_ee; }

void cforacle_set_random_vecs(void *h, const double *vecs, int n_obs) {
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a) {
    agent_t *ag = &m->ee[a];
    if (ag->type != RANDOM_A) continue;
    for (int i = 0; i < n_obs && i < ag->n_state; ++i) {
      const double *v = vecs + ((size_t)a * n_obs + i) * 3;
      ag->random_vecs[i] = v3_make(v[0], v[1], v[2]);
    }
  }
}

void cforacle_get_random_vecs(void *h, double *vecs, int n_obs) {
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a) {
    agent_t *ag = &m->ee[a];
    for (int i = 0; i < n_obs; ++i) {
      double *v = vecs + ((size_t)a * n_obs + i) * 3;
      v3 r = (ag->type == RANDOM_A && i < ag->n_state) ? ag->random_vecs[i] : v3_make(0, 0, 0);
      v[0] = r.x, v[1] = r.y, v[2] = r.z;
    }
  }
}

/* CfManager::setInitialPosition -> setInitialEEPositions -> CfAgent::setInitalPosition
 * (cf_manager.cpp:226-236, cf_agent.cpp:34-46): the real agent APPENDS, the others restart. */
void cforacle_set_initial_position(void *h, const double *p) {
  mgr_t *m = h;
  v3 pos = v3_make(p[0], p[1], p[2]);
  m->init_pos = pos;
  if (m->real.path) {
    m->real.init_pos = pos;
    path_push(&m->real, pos);
  }
  for (int a = 0; a < m->n_ee; ++a) {
    m->ee[a].init_pos = pos;
    m->ee[a].n_path = 0;
    path_push(&m->ee[a], pos);
  }
}

void cforacle_set_real_position(void *h, const double *p) { /* cf_manager.cpp:216-218 */
  mgr_t *m = h;
  path_push(&m->real, v3_make(p[0], p[1], p[2]));
}

double cforacle_rollout_pooled(void *h, int n_threads) {
  mgr_t *m = h;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int i = 0; i < m->n_ee; ++i)
    rollout(&m->ee[i], m->k_a[i], m->k_c[i], m->k_r[i], m->k_d[i], m->pred_dt, m->max_steps);
  return omp_get_wtime() - t0;
}

double cforacle_rollout_threads(void *h) { return cforacle_rollout_pooled(h, 0); }
void cforacle_start_prediction(void *h) { cforacle_rollout_pooled(h, 0); }
void cforacle_stop_prediction(void *h) { (void)h; }

/* CfManager::evaluateAgents, cf_manager.cpp:293-356 */
int cforacle_evaluate_agents(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                             const double *obs_rad, double k_goal_dist, double k_path_len,
                             double k_safe_dist, double k_workspace, const double *ws) {
  (void)n_obs, (void)obs_pos, (void)obs_vel, (void)obs_rad; /* unused by the reference too */
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a) {
    agent_t *ag = &m->ee[a];
    double cost = 0;
    for (int k = 0; k < ag->n_path; ++k) {
      v3 q = ag->path[k];
      double t;
      if (q.x > ws[0]) {
        t = fabs(q.x - ws[0]) * k_workspace, cost += t * t;
      } else if (q.x < ws[1]) {
        t = fabs(q.x - ws[1]) * k_workspace, cost += t * t;
      }
      if (q.y > ws[2]) {
        t = fabs(q.y - ws[2]) * k_workspace, cost += t * t;
      } else if (q.y < ws[3]) {
        t = fabs(q.y - ws[3]) * k_workspace, cost += t * t;
      }
      if (q.z > ws[4]) {
        t = fabs(q.z - ws[4]) * k_workspace, cost += t * t;
      } else if (q.z < ws[5]) {
        t = fabs(q.z - ws[5]) * k_workspace, cost += t * t;
      }
    }
    double goal_dist = dist_from_goal(ag);
    if (goal_dist > m->approach_dist) cost += goal_dist * k_goal_dist;
    double len = 0; /* getPathLength, cf_agent.cpp:26-32 */
    for (int k = 0; k + 1 < ag->n_path; ++k) len += v3_norm(v3_sub(ag->path[k + 1], ag->path[k]));
    cost += len * k_path_len;
    cost += k_safe_dist / ag->min_obs_dist;
    if (ag->min_obs_dist < 2e-5) cost += 10000.0;
    m->costs[a] = cost;
  }
  int min_idx = 0;
  double min_cost = 1.7976931348623157e308; /* numeric_limits<double>::max() */
  for (int a = 0; a < m->n_ee; ++a) {
    if (m->costs[a] < min_cost) {
      min_cost = m->costs[a];
      min_idx = a;
    }
  }
  int take = 1;
  if (m->best.present) { /* hysteresis, :344-350; the reference indexes out of range if id-1 >= A */
    int inc = m->best.id - 1;
    if (inc < m->n_ee && !(m->costs[min_idx] < 0.9 * m->costs[inc])) {
      take = 0;
      min_idx = inc;
    }
  }
  if (take) { /* best_agent_ = ee_agents_[min]->makeCopy() */
    agent_t *ag = &m->ee[min_idx];
    m->best.present = 1, m->best.id = ag->id, m->best.type = ag->type, m->best.n_state = ag->n_state;
    free(m->best.random_vecs);
    m->best.random_vecs = NULL;
    if (ag->type == RANDOM_A) {
      m->best.random_vecs = malloc(sizeof(v3) * (ag->n_state ? ag->n_state : 1));
      memcpy(m->best.random_vecs, ag->random_vecs, sizeof(v3) * ag->n_state);
    }
  }
  return min_idx;
}

/* CfManager::moveRealEEAgent -> RealCfAgent::cfPlanner, cf_manager.cpp:257-263, cf_agent.cpp:343-366 */
void cforacle_move_real_agent(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                              const double *obs_rad, double delta_t, int steps, int agent_id) {
  mgr_t *m = h;
  if (!m->best.present) return; /* the reference dereferences a null best_agent_ here */
  obstacles_t obs = obstacles_from(n_obs, obs_pos, obs_vel, obs_rad);
  agent_t *a = &m->real;
  for (int s = 0; s < steps; ++s) {
    a->force = v3_make(0.0, 0.0, 0.0);
    double k_goal_scale = 1.0;
    if (field_gate_open(a)) {
      circ_force(a, &obs, m->k_c[agent_id], m->best.type, m->best.random_vecs, 1);
      if (v3_norm(a->force) > 1e-5) k_goal_scale = attractor_force_scaling(a, &obs);
    }
    repel_force(a, &obs, m->k_r[agent_id]);
    attractor_force(a, m->k_a[agent_id], m->k_d[agent_id], k_goal_scale);
    update_position_and_velocity(a, delta_t);
  }
  obstacles_free(&obs);
}

/* CfManager::resetEEAgents, cf_manager.cpp:246-255; setVelocity cf_agent.cpp:54-61; setObstacles :63-70 */
void cforacle_reset_agents(void *h, const double *pos, const double *vel, int n_obs, const double *obs_pos,
                           const double *obs_vel, const double *obs_rad) {
  (void)obs_rad; /* radii keep their init() values: setObstacles copies pos and vel only */
  mgr_t *m = h;
  v3 p = v3_make(pos[0], pos[1], pos[2]), v = v3_make(vel[0], vel[1], vel[2]);
  for (int a = 0; a < m->n_ee; ++a) {
    agent_t *ag = &m->ee[a];
    ag->n_path = 0;
    path_push(ag, p);
    double vn = v3_norm(v);
    ag->vel = vn > ag->vel_max ? v3_scale(v, ag->vel_max / vn) : v;
    for (int i = 0; i < n_obs && i < ag->obs.n; ++i) {
      ag->obs.pos[i] = v3_make(obs_pos[3 * i], obs_pos[3 * i + 1], obs_pos[3 * i + 2]);
      ag->obs.vel[i] = v3_make(obs_vel[3 * i], obs_vel[3 * i + 1], obs_vel[3 * i + 2]);
      ag->known[i] = m->real.known[i];
    }
    ag->min_obs_dist = ag->shell;
  }
}

static void put3(double *o, v3 v) { o[0] = v.x, o[1] = v.y, o[2] = v.z; }

void cforacle_get_next_position(void *h, double *p) { put3(p, latest(&((mgr_t *)h)->real)); }
void cforacle_get_next_velocity(void *h, double *p) { put3(p, ((mgr_t *)h)->real.vel); }
void cforacle_get_ee_force(void *h, double *p) { put3(p, ((mgr_t *)h)->real.force); }
double cforacle_get_dist_from_goal(void *h) {
  mgr_t *m = h;
  return v3_norm(v3_sub(m->goal, latest(&m->real))); /* cf_manager.h:87-89 */
}
int cforacle_get_best_agent_type(void *h) {
  mgr_t *m = h;
  return m->best.present ? m->best.type : -1;
}
int cforacle_get_best_agent_id(void *h) {
  mgr_t *m = h;
  return m->best.present ? m->best.id : 0;
}
int cforacle_get_num_prediction_steps(void *h, int agent) { return ((mgr_t *)h)->ee[agent].n_path; }
int cforacle_get_real_num_steps(void *h) { return ((mgr_t *)h)->real.n_path; }

void cforacle_get_agent_summaries(void *h, int *steps, double *length, double *min_obs_dist, int *reached,
                                  double *pred_time_ns, int *agent_type) {
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a) {
    agent_t *ag = &m->ee[a];
    if (steps) steps[a] = ag->n_path;
    if (length) {
      double len = 0;
      for (int k = 0; k + 1 < ag->n_path; ++k) len += v3_norm(v3_sub(ag->path[k + 1], ag->path[k]));
      length[a] = len;
    }
    if (min_obs_dist) min_obs_dist[a] = ag->min_obs_dist;
    if (reached) reached[a] = ag->reached_goal;
    if (pred_time_ns) pred_time_ns[a] = ag->prediction_time;
    if (agent_type) agent_type[a] = ag->type;
  }
}

void cforacle_get_predicted_paths(void *h, double *out, int stride) {
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a)
    for (int k = 0; k < m->ee[a].n_path && k < stride; ++k) put3(out + ((size_t)a * stride + k) * 3, m->ee[a].path[k]);
}

void cforacle_get_agent_velocities(void *h, double *out) {
  mgr_t *m = h;
  for (int a = 0; a < m->n_ee; ++a) put3(out + 3 * a, m->ee[a].vel);
}

int cforacle_get_planned_trajectory(void *h, double *out, int max_points) {
  mgr_t *m = h;
  for (int k = 0; k < m->real.n_path && k < max_points; ++k) put3(out + 3 * k, m->real.path[k]);
  return m->real.n_path;
}

void cforacle_get_obstacle_state(void *h, int n_obs, int *known, double *rot) {
  mgr_t *m = h;
  for (int a = 0; a <= m->n_ee; ++a) {
    const agent_t *ag = a < m->n_ee ? &m->ee[a] : &m->real;
    for (int i = 0; i < n_obs; ++i) {
      size_t k = (size_t)a * n_obs + i;
      int have = i < ag->n_state;
      if (known) known[k] = have && ag->known[i];
      if (rot) put3(rot + 3 * k, have ? ag->rot[i] : v3_make(0, 0, 0));
    }
  }
}

int cforacle_host_threads(void) { return omp_get_max_threads(); }

void cforacle_get_costs(void *h, double *costs) {
  mgr_t *m = h;
  memcpy(costs, m->costs, sizeof(double) * m->n_ee);
}
