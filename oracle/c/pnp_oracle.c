/*
 * ORACLE (test infrastructure; never linked into the product library).
 *
 * Plain-C FP64 restatement of the reference hot path, used (a) as a second, independent
 * checker next to the NumPy restatement and (b) as the multi-threaded "CPU path" baseline of
 * bench.py (cpu_baseline.kind = "port").  Follows, line by line:
 *
 *   IK loop      /root/reference/panda_mujoco_gym/skills/ik_solver.py:50-101
 *   reward       /root/reference/panda_mujoco_gym/envs/panda_env.py:205-245
 *   _is_success  panda_env.py:303-306        goal_distance  panda_env.py:311-315
 *
 * The MuJoCo engine calls (mj_kinematics / mj_jacSite, mujoco==2.3.3, third-party and absent
 * from /root/reference) are restated in MuJoCo's own quaternion formulation over the raw body
 * chain (body_pos, body_quat, jnt_axis, jnt_pos) - NOT the product's canonical z-hinge form -
 * so that the two implementations stay independent.
 *
 * It omits mj_forward's collision / constraint stages, which do not influence the IK result:
 * this baseline is therefore FASTER than the real reference (a conservative baseline).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so the reward
 * arithmetic is bit-identical to the NumPy/Python float64 evaluation order).
 * Pinning: tests/test_oracle.py checks this file against the oracle .py files and against
 * tests/golden/ (reference-generated vectors); "parity unpinned" w.r.t. real MuJoCo beyond
 * the home_wpt constant.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_CHAIN 16
#define NARM 7

typedef struct {
  int32_t nbody;                         /* bodies from the world's child down to the site's body */
  int32_t has_joint[ORACLE_MAX_CHAIN];   /* 1 when the body carries one hinge joint */
  double body_pos[ORACLE_MAX_CHAIN][3];
  double body_quat[ORACLE_MAX_CHAIN][4]; /* wxyz, normalised */
  double jnt_axis[ORACLE_MAX_CHAIN][3];
  double jnt_pos[ORACLE_MAX_CHAIN][3];
  double qpos0[ORACLE_MAX_CHAIN];
  double site_pos[3];
  double site_quat[4];
  double lower[NARM], upper[NARM];
} OracleChain;

typedef struct {
  int32_t max_iters;
  double pos_thresh, damping, step_limit;
} OracleIkParams;

typedef struct {
  int32_t sparse;        /* reward_type == "sparse" */
  int32_t n_tasks;       /* len(task_sequence) */
  double initial_object_height;
  double distance_threshold;
  double high_pick_z;
} OracleRewardParams;

/* ------------------------------------------------------------------ mju_* helpers */
static void mul_quat(double r[4], const double a[4], const double b[4]) {
  double t[4];
  t[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  t[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  t[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  t[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  memcpy(r, t, sizeof t);
}

static void quat2mat(double m[9], const double q[4]) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2 * (q12 - q03);       m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03);       m[4] = q00 - q11 + q22 - q33; m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02);       m[7] = 2 * (q23 + q01);       m[8] = q00 - q11 - q22 + q33;
}

static void rot_vec(double r[3], const double m[9], const double v[3]) {
  double t0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double t1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double t2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}

static void normalize4(double q[4]) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-15) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

/* ------------------------------------------------------------------ mj_kinematics + mj_jacSite
 * Site world position / matrix and (optionally) the 6x7 Jacobian (rows 0-2 jacp, 3-5 jacr). */
static void chain_fk(const OracleChain* c, const double q[NARM], double site_xpos[3],
                     double site_xmat[9], double anchors[NARM][3], double axes[NARM][3]) {
  double xpos[3] = {0, 0, 0}, xquat[4] = {1, 0, 0, 0}, xmat[9];
  int jid = 0;
  quat2mat(xmat, xquat);
  for (int b = 0; b < c->nbody; ++b) {
    double v[3];
    rot_vec(v, xmat, c->body_pos[b]);
    xpos[0] += v[0]; xpos[1] += v[1]; xpos[2] += v[2];
    mul_quat(xquat, xquat, c->body_quat[b]);
    if (c->has_joint[b]) {
      double m0[9], xanchor[3], qloc[4], m1[9], off[3];
      quat2mat(m0, xquat);
      rot_vec(xanchor, m0, c->jnt_pos[b]);
      xanchor[0] += xpos[0]; xanchor[1] += xpos[1]; xanchor[2] += xpos[2];
      rot_vec(axes[jid], m0, c->jnt_axis[b]);
      memcpy(anchors[jid], xanchor, sizeof xanchor);
      double ang = q[jid] - c->qpos0[b];
      double s = sin(0.5 * ang);
      qloc[0] = cos(0.5 * ang);
      qloc[1] = c->jnt_axis[b][0] * s; qloc[2] = c->jnt_axis[b][1] * s; qloc[3] = c->jnt_axis[b][2] * s;
      mul_quat(xquat, xquat, qloc);
      quat2mat(m1, xquat);
      rot_vec(off, m1, c->jnt_pos[b]);
      xpos[0] = xanchor[0] - off[0]; xpos[1] = xanchor[1] - off[1]; xpos[2] = xanchor[2] - off[2];
      ++jid;
    }
    normalize4(xquat);
    quat2mat(xmat, xquat);
  }
  double v[3], sm[9];
  rot_vec(v, xmat, c->site_pos);
  site_xpos[0] = xpos[0] + v[0]; site_xpos[1] = xpos[1] + v[1]; site_xpos[2] = xpos[2] + v[2];
  quat2mat(sm, c->site_quat);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      site_xmat[3 * i + j] = xmat[3 * i] * sm[j] + xmat[3 * i + 1] * sm[3 + j] + xmat[3 * i + 2] * sm[6 + j];
}

static void chain_jac(const double point[3], double anchors[NARM][3], double axes[NARM][3],
                      double jac[6][NARM]) {
  for (int j = 0; j < NARM; ++j) {
    double r[3] = {point[0] - anchors[j][0], point[1] - anchors[j][1], point[2] - anchors[j][2]};
    const double* a = axes[j];
    jac[0][j] = a[1] * r[2] - a[2] * r[1];
    jac[1][j] = a[2] * r[0] - a[0] * r[2];
    jac[2][j] = a[0] * r[1] - a[1] * r[0];
    jac[3][j] = a[0]; jac[4][j] = a[1]; jac[5][j] = a[2];
  }
}

/* 3x3 LU with partial pivoting (what LAPACK dgesv does for np.linalg.solve, ik_solver.py:79) */
static void solve3(double A[3][3], double b[3]) {
  int piv[3] = {0, 1, 2};
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int i = k + 1; i < 3; ++i)
      if (fabs(A[piv[i]][k]) > fabs(A[piv[p]][k])) p = i;
    int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
    for (int i = k + 1; i < 3; ++i) {
      double l = A[piv[i]][k] / A[piv[k]][k];
      A[piv[i]][k] = l;
      for (int j = k + 1; j < 3; ++j) A[piv[i]][j] -= l * A[piv[k]][j];
    }
  }
  double y[3];
  for (int i = 0; i < 3; ++i) {
    y[i] = b[piv[i]];
    for (int j = 0; j < i; ++j) y[i] -= A[piv[i]][j] * y[j];
  }
  for (int i = 2; i >= 0; --i) {
    for (int j = i + 1; j < 3; ++j) y[i] -= A[piv[i]][j] * y[j];
    y[i] /= A[piv[i]][i];
  }
  b[0] = y[0]; b[1] = y[1]; b[2] = y[2];
}

static double norm3(const double v[3]) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* ------------------------------------------------------------------ ik_solver.py:50-101 */
static void ik_solve_one(const OracleChain* c, const OracleIkParams* p, const double target[3],
                         const double q_init[NARM], double q[NARM], double final_pos[3],
                         double* pos_error, int32_t* iterations, uint8_t* flags) {
  double anchors[NARM][3], axes[NARM][3], xmat[9], curr[3];
  memcpy(q, q_init, NARM * sizeof(double));                                   /* :50 */
  chain_fk(c, q, curr, xmat, anchors, axes);                                  /* :51-52 */
  int converged = 0, iters = 0;
  for (int i = 0; i < p->max_iters; ++i) {                                    /* :57 */
    /* :58 mj_kinematics on unchanged qpos == the state we already hold */
    double err[3] = {target[0] - curr[0], target[1] - curr[1], target[2] - curr[2]}; /* :60 */
    double n = norm3(err);                                                    /* :61 */
    if (n < p->pos_thresh) { converged = 1; iters = i + 1; break; }           /* :64-67 */
    double jac[6][NARM];
    chain_jac(curr, anchors, axes, jac);                                      /* :70-72 */
    double A[3][3];
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        double acc = 0;
        for (int j = 0; j < NARM; ++j) acc += jac[r][j] * jac[s][j];
        A[r][s] = acc + (r == s ? p->damping : 0.0);                          /* :79 */
      }
    double y[3] = {err[0], err[1], err[2]};
    solve3(A, y);
    for (int j = 0; j < NARM; ++j) {
      double dq = (jac[0][j] * y[0] + jac[1][j] * y[1]) + jac[2][j] * y[2];   /* :79 J^T y */
      dq = fmin(fmax(dq, -p->step_limit), p->step_limit);                     /* :80 */
      q[j] = fmin(fmax(q[j] + dq, c->lower[j]), c->upper[j]);                 /* :81 */
    }
    chain_fk(c, q, curr, xmat, anchors, axes);                                /* :82-83 */
    iters = i + 1;                                                            /* :85 */
  }
  memcpy(final_pos, curr, sizeof curr);                                       /* :88 */
  double diff[3] = {curr[0] - target[0], curr[1] - target[1], curr[2] - target[2]};
  double fe = norm3(diff);                                                    /* :89 */
  int success = converged && fe < p->pos_thresh * 2;                          /* :92 */
  *pos_error = fe;
  *iterations = iters;
  *flags = (uint8_t)((converged ? 1 : 0) | (success ? 2 : 0));
}

/* ------------------------------------------------------------------ panda_env.py:205-245 */
static const double VERTICAL_QUAT[4] = {1.0, 0.0, -0.0, 0.0};
/* euler2quat([-pi/2, 0, 0]) evaluated in float64 (panda_env.py:30) */
static const double HORIZONTAL_QUAT[4] = {0.7071067811865476, -0.7071067811865475, 0.0, 0.0};

static float reward_one(const OracleRewardParams* p, const double ag[3], const double dg[3],
                        const double ee[3], const double eq[4], double width, int32_t task_idx,
                        float* success) {
  double a[3] = {ee[0] - ag[0], ee[1] - ag[1], ee[2] - ag[2]};
  double b[3] = {ag[0] - dg[0], ag[1] - dg[1], ag[2] - dg[2]};
  double d_reach = norm3(a);                                                  /* :211 */
  double d_place = norm3(b);                                                  /* :212 */
  int gripped = (width < 0.045) && (d_reach < 0.05);                          /* :214-216 */
  int lifted = gripped && (ag[2] - p->initial_object_height > 0.04);          /* :219 */
  int placed = d_place < p->distance_threshold;                               /* :220 */
  if (success) *success = placed ? 1.0f : 0.0f;                               /* :303-306 */
  const double* need = ag[2] > p->high_pick_z ? HORIZONTAL_QUAT : VERTICAL_QUAT; /* :223 */
  double dot = ((eq[0] * need[0] + eq[1] * need[1]) + eq[2] * need[2]) + eq[3] * need[3];
  double ori_err = 1.0 - fabs(dot);                                           /* :224 */
  if (p->sparse) return (float)(-(double)(!placed));                          /* :227-228 */
  double r = -0.003;                                                          /* :231 */
  r += -(0.05 < d_reach ? 0.05 : d_reach); /* :232 Python min(a,b): b only if b < a */
  if (gripped) { r += 2.0; r += (1.0 - ori_err); }                            /* :234-236 */
  if (lifted) r += 4.0;                                                       /* :238-239 */
  if (placed) r += 10.0;                                                      /* :241-242 */
  r += 0.5 * ((double)task_idx / (double)p->n_tasks);                         /* :244 */
  return (float)r;                                                            /* :245 */
}

/* ------------------------------------------------------------------ skills/move.py:76-191
 * MoveIKSkill.reset trajectory planner (adaptive step, accept rule, double failure increment,
 * fallback strategies 1-3, final-point append).  max_outer bounds the while loop: the reference
 * has no such bound and spins forever on unreachable targets (fallback 1 keeps succeeding with
 * ever smaller steps without advancing point_count); 0 = unbounded like the reference. */
typedef struct {
  double pos_thresh;      /* 0.01 */
  int32_t max_traj_points; /* 200 */
  double step_size;       /* 0.01 */
  int32_t max_outer;
  int32_t traj_cap;       /* capacity of the output trajectory (points) */
} OracleMoveParams;

static void move_plan_one(const OracleChain* c, const OracleIkParams* ikp, const OracleMoveParams* mp,
                          const double q_start[NARM], const double target[3], double* traj /*[cap][3]*/,
                          int32_t* traj_len, double q_final[NARM], int32_t* n_solves, int32_t* status) {
  double anchors[NARM][3], axes[NARM][3], xmat[9], pos[3], q[NARM];
  memcpy(q, q_start, sizeof q);
  chain_fk(c, q, pos, xmat, anchors, axes);                                   /* :91 start_pos */
  int len = 0, solves = 0, st = 0;
#define APPEND(P) do { if (len < mp->traj_cap) memcpy(traj + 3 * len, (P), 3 * sizeof(double)); else st |= 4; ++len; } while (0)
  APPEND(pos);                                                                /* :98 */
  int point_count = 0, cf = 0, outer = 0;
  for (;;) {
    double d[3] = {pos[0] - target[0], pos[1] - target[1], pos[2] - target[2]};
    if (!(norm3(d) > mp->pos_thresh && point_count < mp->max_traj_points)) break;   /* :106-107 */
    if (mp->max_outer > 0 && outer >= mp->max_outer) { st |= 2; break; }
    ++outer;
    double dir[3] = {target[0] - pos[0], target[1] - pos[1], target[2] - pos[2]};   /* :110 */
    double dist = norm3(dir);                                                 /* :111 */
    double step = fmin(mp->step_size, dist * 0.1);                            /* :114 */
    step = fmin(step, 0.02);                                                  /* :117 */
    if (cf > 0) step *= 0.5;                                                  /* :118-119 */
    double next[3];
    for (int k = 0; k < 3; ++k) next[k] = dist > step ? pos[k] + dir[k] * step / dist : target[k]; /* :122-125 */
    double qs[NARM], fp[3], err; int32_t it; uint8_t fl;
    ik_solve_one(c, ikp, next, q, qs, fp, &err, &it, &fl); ++solves;          /* :128 */
    if ((fl & 2) && err < mp->step_size * 2) {                                /* :131 */
      APPEND(fp); memcpy(pos, fp, sizeof fp); memcpy(q, qs, sizeof qs); cf = 0;
    } else {
      ++cf;                                                                   /* :142 */
      if (cf >= 3) {                                                          /* :144 */
        double smaller = step * 0.1;                                          /* :149 */
        if (dist > smaller) {
          double fb[3];
          for (int k = 0; k < 3; ++k) fb[k] = pos[k] + dir[k] * smaller / dist;
          ik_solve_one(c, ikp, fb, q, qs, fp, &err, &it, &fl); ++solves;      /* :152 */
          if (fl & 2) { APPEND(fp); memcpy(pos, fp, sizeof fp); memcpy(q, qs, sizeof qs); cf = 0; continue; }
        }
        double alt[3] = {dir[0], 0.0, dir[2]};                                /* :163-164 */
        double an = norm3(alt);
        if (an > 0.001) {
          double ap[3];
          for (int k = 0; k < 3; ++k) ap[k] = pos[k] + (alt[k] / an) * step;
          ik_solve_one(c, ikp, ap, q, qs, fp, &err, &it, &fl); ++solves;      /* :168 */
          if (fl & 2) { APPEND(fp); memcpy(pos, fp, sizeof fp); memcpy(q, qs, sizeof qs); cf = 0; continue; }
        }
        st |= 1;                                                              /* :178-180 break */
        break;
      } else {
        ++cf;                                                                 /* :183 */
        continue;
      }
    }
    ++point_count;                                                            /* :186 */
  }
  double d[3] = {pos[0] - target[0], pos[1] - target[1], pos[2] - target[2]};
  if (norm3(d) > mp->pos_thresh) APPEND(target);                              /* :189-191 */
#undef APPEND
  *traj_len = len; *n_solves = solves; *status = st;
  memcpy(q_final, q, sizeof q);
}

/* ------------------------------------------------------------------ threading */
typedef struct {
  int kind; /* 0 ik, 1 reward, 2 fk, 3 move plan */
  const OracleMoveParams* mvp; const double *mv_q, *mv_goal; double *mv_traj, *mv_qf;
  int32_t *mv_len, *mv_solves, *mv_status;
  int64_t begin, end;
  const OracleChain* chain;
  const OracleIkParams* ikp;
  const OracleRewardParams* rwp;
  const double *targets, *q_init; int64_t q_init_stride;
  double *q_out, *final_pos, *pos_err; int32_t* iters; uint8_t* flags;
  const double *ag, *dg, *ee, *eq, *width; const int32_t* task_idx;
  float *reward, *success;
  const double* fk_q; double *fk_pos, *fk_mat, *fk_jac;
} Job;

static void* worker(void* arg) {
  Job* j = (Job*)arg;
  for (int64_t i = j->begin; i < j->end; ++i) {
    if (j->kind == 0) {
      ik_solve_one(j->chain, j->ikp, j->targets + 3 * i, j->q_init + j->q_init_stride * i,
                   j->q_out + NARM * i, j->final_pos + 3 * i, j->pos_err + i, j->iters + i, j->flags + i);
    } else if (j->kind == 1) {
      j->reward[i] = reward_one(j->rwp, j->ag + 3 * i, j->dg + 3 * i, j->ee + 3 * i, j->eq + 4 * i,
                                j->width[i], j->task_idx[i], j->success ? j->success + i : NULL);
    } else if (j->kind == 3) {
      move_plan_one(j->chain, j->ikp, j->mvp, j->mv_q + NARM * i, j->mv_goal + 3 * i,
                    j->mv_traj + (int64_t)3 * j->mvp->traj_cap * i, j->mv_len + i, j->mv_qf + NARM * i,
                    j->mv_solves + i, j->mv_status + i);
    } else {
      double anchors[NARM][3], axes[NARM][3], jac[6][NARM];
      chain_fk(j->chain, j->fk_q + NARM * i, j->fk_pos + 3 * i, j->fk_mat + 9 * i, anchors, axes);
      if (j->fk_jac) {
        chain_jac(j->fk_pos + 3 * i, anchors, axes, jac);
        memcpy(j->fk_jac + 42 * i, jac, sizeof jac);
      }
    }
  }
  return NULL;
}

static void run_jobs(Job* proto, int64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
  pthread_t tid[256];
  Job jobs[256];
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = *proto;
    jobs[t].begin = t * chunk < n ? t * chunk : n;
    jobs[t].end = (t + 1) * chunk < n ? (t + 1) * chunk : n;
  }
  if (nthreads == 1) { worker(&jobs[0]); return; }
  for (int t = 0; t < nthreads; ++t) pthread_create(&tid[t], NULL, worker, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------------ exported entry points */
void oracle_fk_jac(const OracleChain* chain, const double* q, int64_t n, double* pos, double* mat,
                   double* jac6x7, int nthreads) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 2; j.chain = chain; j.fk_q = q; j.fk_pos = pos; j.fk_mat = mat; j.fk_jac = jac6x7;
  run_jobs(&j, n, nthreads);
}

void oracle_ik_solve(const OracleChain* chain, const OracleIkParams* p, const double* targets,
                     const double* q_init, int64_t q_init_stride, int64_t n, double* q_out,
                     double* final_pos, double* pos_err, int32_t* iters, uint8_t* flags, int nthreads) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 0; j.chain = chain; j.ikp = p; j.targets = targets; j.q_init = q_init;
  j.q_init_stride = q_init_stride; j.q_out = q_out; j.final_pos = final_pos; j.pos_err = pos_err;
  j.iters = iters; j.flags = flags;
  run_jobs(&j, n, nthreads);
}

void oracle_reward(const OracleRewardParams* p, const double* ag, const double* dg, const double* ee,
                   const double* eq, const double* width, const int32_t* task_idx, int64_t n,
                   float* reward, float* success, int nthreads) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 1; j.rwp = p; j.ag = ag; j.dg = dg; j.ee = ee; j.eq = eq; j.width = width;
  j.task_idx = task_idx; j.reward = reward; j.success = success;
  run_jobs(&j, n, nthreads);
}

void oracle_move_plan(const OracleChain* chain, const OracleIkParams* ikp, const OracleMoveParams* mp,
                      const double* q_start, const double* goal, int64_t n, double* traj, int32_t* traj_len,
                      double* q_final, int32_t* n_solves, int32_t* status, int nthreads) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 3; j.chain = chain; j.ikp = ikp; j.mvp = mp; j.mv_q = q_start; j.mv_goal = goal; j.mv_traj = traj;
  j.mv_len = traj_len; j.mv_qf = q_final; j.mv_solves = n_solves; j.mv_status = status;
  run_jobs(&j, n, nthreads);
}

int oracle_abi_version(void) { return 2; }
