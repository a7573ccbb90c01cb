// TEST-ONLY: runs the product's step logic (agent_step / field_pass from
// predictive-multi-agent-framework_b200/csrc) on the HOST with the single-lane HostGroup policy,
// so that the operation order of the CUDA source can be checked against the oracle without a
// GPU. It is compiled by tests/test_host_step.py into tests/_build/ and is not part of libpmaf.so.
#include <cstring>
#include <vector>

#include "../predictive-multi-agent-framework_b200/csrc/pmaf_rollout.cuh"

using namespace pmaf;

extern "C" int hoststep_rollout(int n_obs, const double *obs_pos, const double *obs_vel, const double *obs_rad,
                                const double *goal, double shell, double mass, double rad, double vmax,
                                double approach, double dt, int H, int type, double k_attr, double k_circ,
                                double k_repel, double k_damp, const double *init_pos, const double *p0,
                                const double *v0, double min_obs0, unsigned char *known_io, double *rot_io,
                                const double *random_vecs, float margin, double *path, int *n_path_io,
                                double *v_out, double *min_obs_out, double *path_len_out) {
  StepEnv P;
  P.goal = ld3(goal), P.n_obs = n_obs, P.pred_dt = dt;
  bool dynamic = false;
  for (int i = 0; i < 3 * n_obs; ++i) dynamic = dynamic || obs_vel[i] != 0.0;
  std::vector<double> px(n_obs), py(n_obs), pz(n_obs), rs(n_obs), vx(n_obs), vy(n_obs), vz(n_obs), dx(n_obs),
      dy(n_obs), dz(n_obs);
  std::vector<float4> bp(n_obs);
  for (int i = 0; i < n_obs; ++i) {
    px[i] = obs_pos[3 * i], py[i] = obs_pos[3 * i + 1], pz[i] = obs_pos[3 * i + 2];
    vx[i] = obs_vel[3 * i], vy[i] = obs_vel[3 * i + 1], vz[i] = obs_vel[3 * i + 2];
    dx[i] = vx[i] * dt, dy[i] = vy[i] * dt, dz[i] = vz[i] * dt;
    rs[i] = rad + obs_rad[i];
    bp[i] = broad_phase_record(mk3(px[i], py[i], pz[i]), shell, rs[i], margin);
  }
  SmemObstacles obs;
  obs.px = px.data(), obs.py = py.data(), obs.pz = pz.data(), obs.rs = rs.data();
  obs.vx = vx.data(), obs.vy = vy.data(), obs.vz = vz.data(), obs.dynamic = dynamic;
  std::vector<uint16_t> cand(n_obs + 8);
  double fbuf[3];
  std::vector<uint32_t> words((n_obs + 31) / 32, 0u);
  for (int i = 0; i < n_obs; ++i)
    if (known_io[i]) words[i >> 5] |= 1u << (i & 31);
  KnownBits known;
  known.w = words.data();
  HostGroup g;
  const AgentConsts k = make_agent_consts(k_attr, k_circ, k_repel, k_damp, shell, vmax, approach, mass, rs[n_obs - 1]);
  v3 p = ld3(p0), v = ld3(v0);
  const v3 gl = ld3(goal), ip = ld3(init_pos);
  double min_obs = min_obs0, path_len = 0.0;
  int n_path = *n_path_io;
  double zseg = 1.0;
  bool has_seg = false;
  for (;;) {
    const v3 goal_vec = sub3(gl, p);
    const Prologue pr = dynamic ? step_prologue<false, true>(g, bp.data(), n_obs - 1, cand.data(), goal_vec, p, v, zseg, has_seg, k)
                                : step_prologue<true, false>(g, bp.data(), n_obs - 1, cand.data(), goal_vec, p, v, zseg, has_seg, k);
    const StepNorms &sn = pr.sn;
    path_len += sn.seg_len;
    has_seg = false;
    if (!(sn.dist_goal > 0.1 && n_path < H)) break;
    const v3 prev = p;
    if (dynamic)
      agent_step<false, true>(g, P, obs, bp.data(), cand.data(), fbuf, known, type, k, ip, rot_io, random_vecs, goal_vec, pr, p, v,
                        min_obs);
    else
      agent_step<true, false>(g, P, obs, bp.data(), cand.data(), fbuf, known, type, k, ip, rot_io, random_vecs, goal_vec, pr, p, v,
                       min_obs);
    { const v3 seg = sub3(p, prev); zseg = dot3(seg, seg); has_seg = true; }
    st3(path + 3 * n_path, p);
    ++n_path;
    if (dynamic) {
      for (int i = 0; i < n_obs; ++i) {
        px[i] = px[i] + dx[i], py[i] = py[i] + dy[i], pz[i] = pz[i] + dz[i];
        bp[i].x = (float)px[i], bp[i].y = (float)py[i], bp[i].z = (float)pz[i];
      }
    }
  }
  for (int i = 0; i < n_obs; ++i) known_io[i] = (words[i >> 5] >> (i & 31)) & 1u;
  *n_path_io = n_path;
  st3(v_out, v);
  *min_obs_out = min_obs;
  *path_len_out = path_len;
  return 0;
}
