/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double precision like the numba/python original) of the MetaMaze hot path:
 *   DDA_2D / maze_view                      metagym/metamaze/envs/ray_caster_utils.py:11-62, 66-209
 *   MazeBase.reset / evaluation_rule        metagym/metamaze/envs/maze_base.py:40-95, 191-202
 *   MazeCoreDiscrete3D turn/move/do_action  metagym/metamaze/envs/maze_discrete_3d.py:39-81, 113-127
 *   MazeCore2D do_action/update_observation metagym/metamaze/envs/maze_2d.py:21-34, 89-121
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * Numeric conventions copied from how numba types the original: all geometry in float64; the three per-column
 * tables (cos_hp, cos_abs, sin_abs) are STORED as float32 (ray_caster_utils.py:82-92) and promoted on use; sin/cos of
 * the heading are float32 libm calls because the heading is a float32 scalar (maze_discrete_3d.py:46); float -> int32
 * conversions truncate toward zero; no fused multiply-add (gcc -ffp-contract=off, numba default fastmath=False).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MO_MAX_N 31
#define MO_MAX_HITS 96

typedef struct {
    int32_t n, task_type /*0 SURVIVAL 1 ESCAPE*/, max_steps, view_grid, res_h, res_v;
    double max_vision, fov, l_focal, text_size;
} mo_cfg;

typedef struct {
    int32_t start[2], goal[2];
    double cell_size, wall_height, agent_height, initial_life, max_life, step_reward, goal_reward;
} mo_task;

typedef struct {
    int32_t gx, gy, ori, steps;
    double life;
    double cur_food[MO_MAX_N * MO_MAX_N];
    int32_t revival[MO_MAX_N * MO_MAX_N];
    int32_t wait[MO_MAX_N * MO_MAX_N];
} mo_env;

typedef struct { double dist; int i, j, side; double strength; } mo_hit;

/* DDA_2D, ray_caster_utils.py:11-62.  cos_ori / sin_ori are the float32 table entries.  transp is row-major [n][n]. */
static void mo_dda(const double pos[2], int i, int j, int n, double cell_size, float cos_ori_f, float sin_ori_f,
                   const int8_t *walls, const double *transp, double max_vision, double *hit_dist_o, int *hit_i_o,
                   int *hit_j_o, int *hit_side_o, mo_hit *hits, int *n_hits_o)
{
    const double cos_ori = (double)cos_ori_f, sin_ori = (double)sin_ori_f;
    const double delta_dist_x = fabs(cos_ori) < 1.0e-6 ? 1.0e+6 : fabs(cell_size / cos_ori);
    const double delta_dist_y = fabs(sin_ori) < 1.0e-6 ? 1.0e+6 : fabs(cell_size / sin_ori);
    const double d_x = cos_ori > 0 ? ((i + 1) * cell_size - pos[0]) : (i * cell_size - pos[0]);
    const double d_y = sin_ori > 0 ? ((j + 1) * cell_size - pos[1]) : (j * cell_size - pos[1]);
    double side_dist_x = fabs(cos_ori) < 1.0e-6 ? 1.0e+6 : d_x / cos_ori;
    double side_dist_y = fabs(sin_ori) < 1.0e-6 ? 1.0e+6 : d_y / sin_ori;
    const int delta_i = cos_ori > 0 ? 1 : -1, delta_j = sin_ori > 0 ? 1 : -1;
    int hit_i = i, hit_j = j, hit_side = 0, nh = 0;
    double hit_dist = 0.0;
    if (transp[hit_i * n + hit_j] > 0.01) {                                   /* :25-29 start cell */
        mo_hit h;
        h.i = hit_i; h.j = hit_j; h.strength = transp[hit_i * n + hit_j];
        if (side_dist_x < side_dist_y) { h.dist = side_dist_x; h.side = 0; } else { h.dist = side_dist_y; h.side = 1; }
        hits[nh++] = h;
    }
    while (hit_dist < max_vision) {
        if (side_dist_x < side_dist_y) {
            hit_i += delta_i;
            side_dist_y -= side_dist_x;
            hit_dist += side_dist_x;
            if (hit_i < 0 || hit_i >= n) {
                if (hit_j < 0 || hit_j >= n) { hit_dist = 1.0e+6; break; }   /* :36-39 (both out) */
            } else if (hit_j >= 0 && hit_j < n) {
                if (transp[hit_i * n + hit_j] > 0.01 && nh < MO_MAX_HITS) {
                    mo_hit h = {hit_dist, hit_i, hit_j, 0, transp[hit_i * n + hit_j]};
                    hits[nh++] = h;
                }
                if (walls[hit_i * n + hit_j] > 0) { hit_side = 0; break; }
            }
            side_dist_x = delta_dist_x;
        } else {
            hit_j += delta_j;
            side_dist_x -= side_dist_y;
            hit_dist += side_dist_y;
            if (hit_i < 0 || hit_i >= n) {
                if (hit_j < 0 || hit_j >= n) { hit_dist = 1.0e+6; break; }
            } else if (hit_j >= 0 && hit_j < n) {
                if (transp[hit_i * n + hit_j] > 0.01 && nh < MO_MAX_HITS) {
                    mo_hit h = {hit_dist, hit_i, hit_j, 1, transp[hit_i * n + hit_j]};
                    hits[nh++] = h;
                }
                if (walls[hit_i * n + hit_j] > 0) { hit_side = 1; break; }
            }
            side_dist_y = delta_dist_y;
        }
    }
    *hit_dist_o = hit_dist; *hit_i_o = hit_i; *hit_j_o = hit_j; *hit_side_o = hit_side; *n_hits_o = nh;
}

static inline int32_t mo_trunc(double x) { return (int32_t)x; }

/* rgb_array[d_h, d_v, :] = light * (alpha * FAR_RGB + (1 - alpha) * texel)   (FAR_RGB = 0) */
static inline void mo_paint(int32_t *px, double light, double alpha, const uint8_t *texel)
{
    for (int c = 0; c < 3; ++c) px[c] = mo_trunc(light * (alpha * 0.0 + (1.0 - alpha) * (double)texel[c]));
}

/* rgb = (1 - tf) * rgb + tf * TRANSPARENT_RGB   (TRANSPARENT_RGB = 0,255,0), ray_caster_utils.py:8 */
static inline void mo_blend(int32_t *px, double tf)
{
    static const double T[3] = {0.0, 255.0, 0.0};
    for (int c = 0; c < 3; ++c) px[c] = mo_trunc((1.0 - tf) * (double)px[c] + tf * T[c]);
}

/*
 * maze_view, ray_caster_utils.py:66-209.
 *   walls, texts int8 [n][n]; transp float64 [n][n]; tex uint8 [n_tex][ts][ts][3]; ceil uint8 [ts][ts][3]
 *   (s_ori, c_ori): sin/cos of the heading as numba computes them -- float32 libm results promoted for the discrete
 *   maze (float32 heading), float64 libm results for the continuous maze (python-float heading); out int32 [res_h][res_v][3]; transp_mark scratch float [res_h][res_v]
 */
void mo_maze_view(const mo_cfg *c, const mo_task *t, const double pos[2], double s_ori, double c_ori,
                  const int8_t *walls, const double *transp, const int8_t *texts, const uint8_t *tex,
                  const uint8_t *ceil_tex, int ts, int32_t *out, float *transp_mark)
{
    const int n = c->n, H = c->res_h, V = c->res_v;
    const double vision_height = t->agent_height, ceil_height = t->wall_height, cell_size = t->cell_size;
    const double text_size = c->text_size, max_vision = c->max_vision, l_focal = c->l_focal;
    const double half_h = tan(c->fov / 2) * l_focal;
    const double half_v = half_h * V / H;
    const double pixel_size = 2.0 * half_h / H;
    const double text_to_cell = text_size / cell_size;
    const double pixel_factor = pixel_size / l_focal;
    float cos_hp_a[4096], cos_abs_a[4096], sin_abs_a[4096];

    memset(out, 0, sizeof(int32_t) * (size_t)H * V * 3);
    memset(transp_mark, 0, sizeof(float) * (size_t)H * V);
    double tan_hp = (-0.5 - (double)H / 2) * pixel_factor;
    for (int d_h = 0; d_h < H; ++d_h) {                                       /* :86-92 */
        tan_hp += pixel_factor;
        const double cos_hp = sqrt(1.0 / (1.0 + tan_hp * tan_hp));
        const double sin_hp = tan_hp * cos_hp;
        sin_abs_a[d_h] = (float)(sin_hp * (double)c_ori + cos_hp * (double)s_ori);
        cos_abs_a[d_h] = (float)(cos_hp * (double)c_ori - sin_hp * (double)s_ori);
        cos_hp_a[d_h] = (float)cos_hp;
    }

    /* floor :95-126 */
    for (int d_v = V - 1; d_v > V / 2; --d_v) {
        const double v_screen = (d_v + 0.5) * pixel_size - half_v;
        const double distance = vision_height / v_screen * l_focal;
        const double light = v_screen / l_focal;
        if (distance > max_vision) continue;
        for (int d_h = 0; d_h < H; ++d_h) {
            const double eff = distance / (double)cos_hp_a[d_h];
            const double alpha = fmin(1.0, fmax(2.0 * eff / max_vision - 1.0, 0.0)) * light;
            const double hit_x = eff * (double)cos_abs_a[d_h] + pos[0];
            const double hit_y = eff * (double)sin_abs_a[d_h] + pos[1];
            double fi = hit_x / cell_size, fj = hit_y / cell_size;
            double d_i = fi - floor(fi), d_j = fj - floor(fj);
            const int i = mo_trunc(fi), j = mo_trunc(fj);
            if (i < n && i >= 0 && j < n && j >= 0) {
                const int text_id = texts[i * n + j];
                d_i /= text_to_cell; d_j /= text_to_cell;
                d_i -= floor(d_i); d_j -= floor(d_j);
                d_i *= ts; d_j *= ts;
                int32_t *px = out + ((size_t)d_h * V + d_v) * 3;
                mo_paint(px, light, alpha, tex + (((size_t)text_id * ts + mo_trunc(d_i)) * ts + mo_trunc(d_j)) * 3);
                if (transp[i * n + j] > 0.01) {
                    mo_blend(px, transp[i * n + j] * 0.50 + 0.10);
                    transp_mark[(size_t)d_h * V + d_v] = 1.0f;
                }
            }
        }
    }
    /* ceiling :129-153 */
    for (int d_v = 0; d_v < V / 2; ++d_v) {
        const double v_screen = half_v - (d_v + 0.5) * pixel_size;
        const double distance = (ceil_height - vision_height) / v_screen * l_focal;
        const double light = v_screen / l_focal;
        if (distance > max_vision) continue;
        for (int d_h = 0; d_h < H; ++d_h) {
            const double eff = distance / (double)cos_hp_a[d_h];
            const double alpha = fmin(1.0, fmax(2.0 * eff / max_vision - 1.0, 0.0));
            const double hit_x = eff * (double)cos_abs_a[d_h] + pos[0];
            const double hit_y = eff * (double)sin_abs_a[d_h] + pos[1];
            const int t_i = mo_trunc(hit_x / cell_size), t_j = mo_trunc(hit_y / cell_size);
            double fi = hit_x / text_size, fj = hit_y / text_size;
            double d_i = fi - floor(fi), d_j = fj - floor(fj);
            d_i *= ts; d_j *= ts;
            int32_t *px = out + ((size_t)d_h * V + d_v) * 3;
            mo_paint(px, light, alpha, ceil_tex + ((size_t)mo_trunc(d_i) * ts + mo_trunc(d_j)) * 3);
            if (t_i >= 0 && t_i < n && t_j >= 0 && t_j < n && transp[t_i * n + t_j] > 0) {
                mo_blend(px, transp[t_i * n + t_j] * 0.50 + 0.10);
                transp_mark[(size_t)d_h * V + d_v] = 1.0f;
            }
        }
    }
    /* walls :156-205 */
    for (int d_h = 0; d_h < H; ++d_h) {
        const int i = mo_trunc(pos[0] / cell_size), j = mo_trunc(pos[1] / cell_size);
        double hit_dist;
        int hit_i, hit_j, hit_side, nh;
        mo_hit hits[MO_MAX_HITS];
        mo_dda(pos, i, j, n, cell_size, cos_abs_a[d_h], sin_abs_a[d_h], walls, transp, max_vision, &hit_dist, &hit_i,
               &hit_j, &hit_side, hits, &nh);
        if (hit_dist > max_vision) continue;
        const double alpha = fmin(1.0, fmax(2.0 * hit_dist / max_vision - 1.0, 0.0));
        const int ci = hit_i < 0 ? 0 : (hit_i >= n ? n - 1 : hit_i), cj = hit_j < 0 ? 0 : (hit_j >= n ? n - 1 : hit_j);
        const int text_id = texts[ci * n + cj];
        const double hit_pt_x = hit_dist * (double)cos_abs_a[d_h] + pos[0];
        const double hit_pt_y = hit_dist * (double)sin_abs_a[d_h] + pos[1];
        double local_h, light;
        if (hit_side == 0) {
            local_h = hit_pt_y / cell_size; local_h -= floor(local_h);
            light = (double)fabsf(cos_abs_a[d_h]);
        } else {
            local_h = hit_pt_x / cell_size; local_h -= floor(local_h);
            light = (double)fabsf(sin_abs_a[d_h]);
        }
        double ratio = hit_dist * (double)cos_hp_a[d_h] / l_focal;
        double top_v = (ceil_height - vision_height) / ratio, bot_v = vision_height / ratio;
        int v_s = mo_trunc((half_v - top_v) / pixel_size), v_e = mo_trunc((half_v + bot_v) / pixel_size);
        if (v_s < 0) v_s = 0;
        if (v_e > V) v_e = V;
        for (int d_v = v_s; d_v < v_e; ++d_v) {
            const double local_v = (half_v - (d_v + 0.5) * pixel_size) * ratio + vision_height;
            double d_i = local_h / text_size, d_j = local_v / text_size;
            d_i -= floor(d_i); d_j -= floor(d_j);
            const int ti = mo_trunc(ts * d_i), tj = mo_trunc(ts * d_j);
            mo_paint(out + ((size_t)d_h * V + d_v) * 3, light, alpha, tex + (((size_t)text_id * ts + ti) * ts + tj) * 3);
        }
        for (int k = 0; k < nh; ++k) {                                         /* :191-205 */
            ratio = hits[k].dist * (double)cos_hp_a[d_h] / l_focal;
            const double tf = hits[k].strength * 0.50 + 0.10;
            top_v = (ceil_height - vision_height) / ratio; bot_v = vision_height / ratio;
            v_s = mo_trunc((half_v - top_v) / pixel_size); v_e = mo_trunc((half_v + bot_v) / pixel_size);
            if (v_s < 0) v_s = 0;
            if (v_e > V) v_e = V;
            for (int d_v = v_s; d_v < v_e; ++d_v)
                if (transp_mark[(size_t)d_h * V + d_v] < 1) mo_blend(out + ((size_t)d_h * V + d_v) * 3, tf);
        }
    }
}

/* python slice bound normalisation for obs[a:b] on an axis of length len */
static int mo_slice_bound(int x, int len)
{
    if (x < 0) { x += len; if (x < 0) x = 0; }
    if (x > len) x = len;
    return x;
}

static void mo_transparents(const mo_cfg *c, const mo_task *t, const mo_env *e, double *transp)
{
    const int n = c->n;
    if (c->task_type == 0) memcpy(transp, e->cur_food, sizeof(double) * n * n);   /* alias, maze_base.py:57 */
    else {
        memset(transp, 0, sizeof(double) * n * n);
        transp[t->goal[0] * n + t->goal[1]] = 1.0;                                /* maze_base.py:59-60 */
    }
}

/* MazeCoreDiscrete3D.update_observation, maze_discrete_3d.py:113-127 */
void mo_observe_3d(const mo_cfg *c, const mo_task *t, const mo_env *e, const int8_t *walls, const int8_t *texts,
                   const uint8_t *tex, const uint8_t *ceil_tex, int ts, int32_t *obs, float *scratch)
{
    const int n = c->n, H = c->res_h, V = c->res_v;
    double transp[MO_MAX_N * MO_MAX_N];
    mo_transparents(c, t, e, transp);
    const double pos[2] = {e->gx * t->cell_size + 0.5 * t->cell_size, e->gy * t->cell_size + 0.5 * t->cell_size};
    static const float CH[4] = {0.0f, 0.5f, 1.0f, 1.5f};
    const float ori = CH[e->ori] * (float)3.1415926;  /* float32 array * weak python float, maze_discrete_3d.py:46 */
    (void)n;
    mo_maze_view(c, t, pos, (double)sinf(ori), (double)cosf(ori), walls, transp, texts, tex, ceil_tex, ts, obs, scratch);
    if (c->task_type == 0) {
        const double lb_sx = 0.10 * V, lb_sy = 0.10 * V, lb_w = 0.05 * H, lb_l = 0.80 * V;   /* :42-45 */
        const double l = e->life / t->max_life * lb_l;
        int sx = mo_slice_bound((int)lb_sx, H), ex = mo_slice_bound((int)(lb_sx + l), H);
        int sy = mo_slice_bound((int)lb_sy, V), ey = mo_slice_bound((int)(lb_sy + lb_w), V);
        for (int x = sx; x < ex; ++x)
            for (int y = sy; y < ey; ++y) {
                int32_t *px = obs + ((size_t)x * V + y) * 3;
                px[0] = 255; px[1] = 0; px[2] = 0;
            }
    }
}

/* MazeCore2D.update_observation, maze_2d.py:89-121 -> float32 [(2g+1)][(2g+1)] */
void mo_observe_2d(const mo_cfg *c, const mo_task *t, const mo_env *e, const int8_t *walls, float *obs)
{
    const int n = c->n, g = c->view_grid, W = 2 * g + 1;
    for (int a = 0; a < W; ++a)
        for (int b = 0; b < W; ++b) {
            const int x = e->gx - g + a, y = e->gy - g + b;
            float v = -1.0f;
            if (x >= 0 && x < n && y >= 0 && y < n) {
                v = (float)(-(int)walls[x * n + y]);
                if (c->task_type == 0) v = (float)((double)v + e->cur_food[x * n + y]);
                else v = (float)((double)v + ((x == t->goal[0] && y == t->goal[1]) ? 1.0 : 0.0));
            }
            obs[a * W + b] = v;
        }
    if (c->task_type == 0) obs[g * W + g] = (float)e->life;
}

/* MazeBase.reset (+ MazeCoreDiscrete3D.reset), maze_base.py:40-63 */
void mo_reset(const mo_cfg *c, const mo_task *t, const double *food, const int32_t *interval, mo_env *e)
{
    const int n = c->n;
    e->gx = t->start[0]; e->gy = t->start[1]; e->ori = 0; e->steps = 0;
    e->life = t->initial_life;
    for (int k = 0; k < n * n; ++k) { e->cur_food[k] = food[k]; e->revival[k] = interval[k]; e->wait[k] = 0; }
}

/* MazeBase.evaluation_rule, maze_base.py:65-95 */
static void mo_evaluate(const mo_cfg *c, const mo_task *t, const double *food, const int32_t *interval, mo_env *e,
                        double *reward, int *done)
{
    const int n = c->n, idx = e->gx * n + e->gy;
    e->steps += 1;
    const int over = e->steps > c->max_steps - 1;                               /* :191-192 */
    if (c->task_type == 0) {
        double r = 0.0;
        if (e->cur_food[idx] > 1.0e-2) { r = e->cur_food[idx]; e->wait[idx] = 1; e->cur_food[idx] = 0.0; }
        e->life += r + t->step_reward;
        e->life = e->life < t->max_life ? e->life : t->max_life;
        *done = (e->life < 0.0) || over;
        for (int k = 0; k < n * n; ++k) {
            e->revival[k] -= e->wait[k];
            if (e->revival[k] < 0) { e->cur_food[k] = food[k]; e->revival[k] = interval[k]; e->wait[k] = 0; }
        }
        *reward = r;
    } else {
        const int goal = (e->gx == t->goal[0] && e->gy == t->goal[1]);
        *reward = t->step_reward + goal * t->goal_reward;
        *done = goal || over;
    }
}

/* MetaMazeDiscrete3D.step -> do_action, maze_env.py:59-75, maze_discrete_3d.py:51-81 */
void mo_step_3d(const mo_cfg *c, const mo_task *t, const int8_t *walls, const double *food, const int32_t *interval,
                mo_env *e, int action, double *reward, int *done)
{
    static const int TURN[4] = {-1, 1, 0, 0}, MOVE[4] = {0, 0, -1, 1};          /* DISCRETE_ACTIONS, maze_env.py:14 */
    const int n = c->n;
    e->ori = ((e->ori + TURN[action]) % 4 + 4) % 4;
    int tx = e->gx, ty = e->gy;
    const int s = MOVE[action];
    if (e->ori == 0) tx += s; else if (e->ori == 1) ty += s; else if (e->ori == 2) tx -= s; else ty -= s;
    if (tx >= 0 && tx < n && ty >= 0 && ty < n && walls[tx * n + ty] == 0) { e->gx = tx; e->gy = ty; }
    mo_evaluate(c, t, food, interval, e, reward, done);
}

/* MetaMaze2D.step -> do_action, maze_env.py:189-206, maze_2d.py:21-34 */
void mo_step_2d(const mo_cfg *c, const mo_task *t, const int8_t *walls, const double *food, const int32_t *interval,
                mo_env *e, int action, double *reward, int *done)
{
    static const int DX[4] = {-1, 1, 0, 0}, DY[4] = {0, 0, -1, 1};
    const int n = c->n;
    int tx = e->gx + DX[action], ty = e->gy + DY[action];
    if (tx < 0) tx += n;   /* numpy negative-index wrap; unreachable with the border walls of maze_task.py:60 */
    if (ty < 0) ty += n;
    if (tx < n && ty < n && walls[tx * n + ty] < 1) { e->gx = tx; e->gy = ty; }
    mo_evaluate(c, t, food, interval, e, reward, done);
}

int mo_env_size(void) { return (int)sizeof(mo_env); }

/* ---------------------------------------------------------------------------------------------------------------
 * MetaMazeContinuous3D: maze_continuous_3d.py:47-56 (do_action), dynamics.py:16-92 (numba + numpy), with the typing
 * the reference gets for float32 actions (what its action_space.sample() yields): turn_rate / walk_speed float32,
 * heading float64, position float32.
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct {
    float pos[2];
    int32_t pos_is_list;   /* 1 until the first step: reset() stores the cell centre as python floats */
    double ori;
} mo_cont;

static float mo_sum2f(float a, float b) { float c = 0.0f; c += a; c += b; return c; }   /* numba np.sum over float32[2] */

/* nearest_point, dynamics.py:16-29.  pos, l1, l2 float32[2] -> distance (float32), nearest point */
static float mo_nearest_point(const float pos[2], const float l1[2], const float l2[2], float np_out[2])
{
    float u[2] = {l2[0] - l1[0], l2[1] - l1[1]};
    const float edge_norm = sqrtf(mo_sum2f(u[0] * u[0], u[1] * u[1]));
    const double den = 1.0e-6 > (double)edge_norm ? 1.0e-6 : (double)edge_norm;
    u[0] = (float)((double)u[0] / den); u[1] = (float)((double)u[1] / den);
    const float dist_1 = mo_sum2f((pos[0] - l1[0]) * u[0], (pos[1] - l1[1]) * u[1]);
    const float *q;
    float lp[2];
    if (dist_1 > edge_norm) q = l2;
    else if (dist_1 < 0) q = l1;
    else { lp[0] = l1[0] + dist_1 * u[0]; lp[1] = l1[1] + dist_1 * u[1]; q = lp; }
    const float a = pos[0] - q[0], b = pos[1] - q[1];
    np_out[0] = q[0]; np_out[1] = q[1];
    return sqrtf(mo_sum2f(a * a, b * b));
}

/* collision_force, dynamics.py:31-56 */
static void mo_collision_force(const float dv[2], double cell_size, double col_dist, float out[2])
{
    static const float O10[2] = {0.5f, 0.5f}, O01[2] = {-0.5f, 0.5f}, Om0[2] = {-0.5f, -0.5f}, O0m[2] = {0.5f, -0.5f};
    double dist = (double)sqrtf(mo_sum2f(dv[0] * dv[0], dv[1] * dv[1]));
    const double eff = col_dist / cell_size;
    out[0] = out[1] = 0.0f;
    if (dist > 0.708 + eff) return;
    if (fabsf(dv[0]) < 0.5f && fabsf(dv[1]) < 0.5f) {
        const float k = (float)(0.50 / (dist > 1.0e-6 ? dist : 1.0e-6) * (0.708 + eff - dist) * cell_size);
        out[0] = k * dv[0]; out[1] = k * dv[1];
        return;
    }
    const int x_pos = (dv[0] + dv[1] > 0), y_pos = (dv[1] - dv[0] > 0);
    float np[2];
    if (x_pos && y_pos) dist = (double)mo_nearest_point(dv, O10, O01, np);
    else if (!x_pos && y_pos) dist = (double)mo_nearest_point(dv, O01, Om0, np);
    else if (!x_pos && !y_pos) dist = (double)mo_nearest_point(dv, Om0, O0m, np);
    else dist = (double)mo_nearest_point(dv, O0m, O10, np);
    if (eff < dist) return;
    float o[2] = {dv[0] - np[0], dv[1] - np[1]};
    const float on = sqrtf(mo_sum2f(o[0] * o[0], o[1] * o[1]));
    const double inv = 1.0 / (1.0e-6 > (double)on ? 1.0e-6 : (double)on);
    o[0] = (float)((double)o[0] * inv); o[1] = (float)((double)o[1] * inv);
    const float k = (float)(0.50 * (eff - dist) * cell_size);
    out[0] = k * o[0]; out[1] = k * o[1];
}

/* vector_move_with_collision, dynamics.py:58-92, for float32 (turn, walk) already clipped/scaled by do_action */
static void mo_vector_move_with_collision(mo_cont *st, float turn_rate, float walk_speed, double deta_t, int n,
                                          const int8_t *walls, double cell_size, double col_dist)
{
    if (walk_speed < 0) walk_speed = walk_speed * 0.5f;
    float tmp[2] = {st->pos[0], st->pos[1]};
    double ori = st->ori;
    const int iters = (int)(100 * deta_t);
    for (int it = 0; it < iters; ++it) {
        /* vector_move(ori, turn_rate, walk_speed, 0.01) */
        double fin = ori + (double)turn_rate * 0.01;
        const double off_ori = 0.5 * (fin + ori);
        const double off = (double)walk_speed * 0.01;
        const float d[2] = {(float)(cos(off_ori) * off), (float)(sin(off_ori) * off)};
        while (fin > 6.2831852) fin -= 6.2831852;
        while (fin < 0) fin += 6.2831852;
        ori = fin;
        const float ex[2] = {tmp[0] + d[0], tmp[1] + d[1]};
        const float cs = (float)cell_size;
        const float ec[2] = {ex[0] / cs, ex[1] / cs};
        float col[2] = {0.0f, 0.0f};
        for (int i = -1; i < 2; ++i)
            for (int j = -1; j < 2; ++j) {
                const int wi = i + (int)ec[0], wj = j + (int)ec[1];
                if (wi > -1 && wi < n && wj > -1 && wj < n && walls[wi * n + wj] > 0) {
                    const float dv[2] = {ec[0] - floorf(ec[0]) - (float)(i + 0.5), ec[1] - floorf(ec[1]) - (float)(j + 0.5)};
                    float f[2];
                    mo_collision_force(dv, cell_size, col_dist, f);
                    col[0] += f[0]; col[1] += f[1];
                }
            }
        tmp[0] = col[0] + ex[0]; tmp[1] = col[1] + ex[1];
    }
    st->ori = ori;
    st->pos[0] = tmp[0]; st->pos[1] = tmp[1];
    st->pos_is_list = 0;
}

void mo_reset_c3d(const mo_cfg *c, const mo_task *t, const double *food, const int32_t *interval, mo_env *e,
                  mo_cont *st)
{
    mo_reset(c, t, food, interval, e);
    st->pos[0] = (float)(t->start[0] * t->cell_size + 0.5 * t->cell_size);
    st->pos[1] = (float)(t->start[1] * t->cell_size + 0.5 * t->cell_size);
    st->pos_is_list = 1;
    st->ori = 0.0;
}

/* MetaMazeContinuous3D.step, maze_env.py:129-146 -> do_action maze_continuous_3d.py:47-56.  tr, ws: float32 actions */
void mo_step_c3d(const mo_cfg *c, const mo_task *t, const int8_t *walls, const double *food, const int32_t *interval,
                 mo_env *e, mo_cont *st, float tr, float ws, double *reward, int *done)
{
    float turn = tr < -1.0f ? -1.0f : (tr > 1.0f ? 1.0f : tr);
    turn = turn * (float)3.1415926;                                /* np.clip(float32) * python float -> float32 */
    const float walk = ws < -1.0f ? -1.0f : (ws > 1.0f ? 1.0f : ws);
    mo_vector_move_with_collision(st, turn, walk, 0.10, c->n, walls, t->cell_size, 0.20);
    const float cs = (float)t->cell_size;                          /* get_loc_grid: float32 / python float */
    e->gx = (int)(st->pos[0] / cs);
    e->gy = (int)(st->pos[1] / cs);
    mo_evaluate(c, t, food, interval, e, reward, done);
}

void mo_observe_c3d(const mo_cfg *c, const mo_task *t, const mo_env *e, const mo_cont *st, const int8_t *walls,
                    const int8_t *texts, const uint8_t *tex, const uint8_t *ceil_tex, int ts, int32_t *obs,
                    float *scratch)
{
    const int H = c->res_h, V = c->res_v;
    double transp[MO_MAX_N * MO_MAX_N];
    mo_transparents(c, t, e, transp);
    const double pos[2] = {(double)st->pos[0], (double)st->pos[1]};
    mo_maze_view(c, t, pos, sin(st->ori), cos(st->ori), walls, transp, texts, tex, ceil_tex, ts, obs, scratch);
    if (c->task_type == 0) {
        const double lb_sx = 0.10 * V, lb_sy = 0.10 * V, lb_w = 0.05 * H, lb_l = 0.80 * V;
        const double l = e->life / t->max_life * lb_l;
        int sx = mo_slice_bound((int)lb_sx, H), ex = mo_slice_bound((int)(lb_sx + l), H);
        int sy = mo_slice_bound((int)lb_sy, V), ey = mo_slice_bound((int)(lb_sy + lb_w), V);
        for (int x = sx; x < ex; ++x)
            for (int y = sy; y < ey; ++y) {
                int32_t *px = obs + ((size_t)x * V + y) * 3;
                px[0] = 255; px[1] = 0; px[2] = 0;
            }
    }
}

int mo_cont_size(void) { return (int)sizeof(mo_cont); }

