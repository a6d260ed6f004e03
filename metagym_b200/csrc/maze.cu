// MetaMaze hot path for sm_100a: 2-D grid step (one thread = one env) and discrete-3D step + first-person raycast
// render (one persistent CTA per SM, textures staged ONCE into shared memory by TMA bulk copy, every env's maze tile
// staged by TMA + mbarrier, pixels leave through double-buffered shared-memory chunks and TMA bulk stores).
//
// Replaces (reference file:line, PaddlePaddle/MetaGym):
//   MazeBase.set_task / reset / evaluation_rule      metagym/metamaze/envs/maze_base.py:19-95, 191-202
//   MazeCore2D.do_action / update_observation        metagym/metamaze/envs/maze_2d.py:21-34, 89-121
//   MazeCoreDiscrete3D.turn / move / do_action / update_observation
//                                                    metagym/metamaze/envs/maze_discrete_3d.py:39-81, 113-127
//   DDA_2D / maze_view                               metagym/metamaze/envs/ray_caster_utils.py:11-62, 66-209
//
// Exactness: grid state, done flags, float64 rewards/life and the rendered integers are bit-identical to the reference.
// The renderer therefore computes pixel geometry in float64 with the reference's operation order, float32 column
// tables, truncating float->int conversions and NO fused multiply-add (this file is compiled with -fmad=false).
// Food respawn is evaluated lazily from per-food "eaten at step s" stamps: a cell eaten at step s is
// edible again at the check of step t iff t > s + interval and visible after step t iff t >= s + interval, which is
// what maze_base.py:74-75,83-88 does with its whole-grid countdown arrays.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "mgb_common.cuh"

namespace {

constexpr int kMaxN = 31;
constexpr int kNever = INT_MIN / 2;
constexpr int kRenderThreads = 512;
constexpr int kGeoThreads = 128;            // direct renderer, pipelined: threads that prepare the next env's records
constexpr int kMaxHitsCap = 48;

struct TaskHdr {                 // 112 bytes, head of every task blob
    int32_t start[2], goal[2];
    double cell_size, wall_height, agent_height, initial_life, max_life, step_reward, goal_reward;
    int32_t n_food;
    int32_t cls;                 // index of the (agent_height, wall_height) class -> pose-independent eff table
    // x / d == x * (1/d) exactly when d is a power of two: the renderer's divisions by cell_size (2.0) and
    // text_size / cell_size (0.5) then cost one multiply instead of a ~40-instruction IEEE division.
    double inv_cell, inv_t2c;
    int32_t cell_pow2, t2c_pow2;
};

struct MazeConst {
    int kind, task_type, n, max_steps, view_grid, res_h, res_v, obs_dtype;
    int n_tex, ts;
    int f_max;                   // food slots per env
    int max_hits;                // transparent crossings kept per column
    int run_px;                  // pixels per warp run (768 B of output): one bulk store per run
    int text_pow2;               // text_size is a power of two
    double inv_text;
    int n_cls;                   // height classes with a precomputed eff table (0 = compute per pixel)
    int pipe;                    // direct renderer: two record sets, geometry of env e + 1 under the pixels of env e
    int hits_in_global;          // large screens: the per-column crossing lists live in a global scratch, not smem
    int blob_bytes;              // bytes of one task blob (multiple of 16)
    int off_walls, off_texts, off_fidx, off_fval, off_fint;   // offsets inside a blob
    double max_vision, l_focal, text_size;
    double half_h, half_v, pixel_size;                        // host-computed like ray_caster_utils.py:68-70
    // life bar (maze_discrete_3d.py:42-45)
    double lb_sx, lb_sy, lb_w, lb_l;
};

struct MazeArgs {
    int64_t n, n_pad, env_base;
    int4 *agent;                 // gx, gy, ori, steps
    double *life;
    int32_t *eaten;              // [f_max][n_pad]
    const int32_t *env2task;
    const uint8_t *blobs;        // [n_tasks][blob_bytes]
    const uint32_t *tex;         // packed 0x00BBGGRR, (n_tex + 1) * ts * ts, ceiling last
    const float *coltab;         // [4][3][res_h]: cos_hp, cos_abs, sin_abs per heading
    const double *efftab;        // [n_cls][res_h * res_v]: distance(d_v) / cos_hp(d_h), pose independent
    const double *fogtab;        // [n_cls][res_h * res_v]: alpha = clip(2 eff / max_vision - 1, 0, 1) of that distance
    // pose cache (memoised static layers, see maze3d_compose_kernel)
    const int4 *poses;           // FILL: [n_slots] task, gx, gy, ori
    const int32_t *pose_index;   // [n_tasks][n*n*4] -> slot or -1
    const void *pose_rec;        // [n_tasks][n*n*4] PoseRec (pose_index + c_fmask + c_vbase merged), nullptr: use the three tables
    uint32_t *c_px;              // [n_slots][H*V]   10-bit R | G<<10 | B<<20 | in_wall<<30
    uint8_t *c_fid;              // [n_slots][H*V]   food slot of the floor/ceiling cell under the pixel, 0xFF none
    uint8_t *c_rgb8;             // [n_slots][H*V*3] finished uint8 pixel with EVERY food of the task present (baked)
    uint32_t *c_px_all;          // int32 mode: [n_slots][H*V] packed like c_px, every food present (baked)
    uint64_t *c_fmask;           // [n_slots][2] food slots that can change this pose's image at all
    // variant frames (uint8 SURVIVAL): a pose whose image depends on k <= kVariantBits foods has all 2^k finished frames
    // baked, so a step never recomputes a pixel -- it picks the frame of the foods currently visible
    const int32_t *c_vbase;      // [n_slots] index of the pose's first extra frame in c_var8, or -1 (all-present frame only)
    uint8_t *c_var8;             // [n_var_frames][H*V*3]; variant v of a pose = frame c_vbase + v, v = the visible foods'
                                 // bits compacted in ascending slot order; the all-visible variant is c_rgb8[slot] itself
    const void *bake_desc;       // bake mode over variant frames: [n_items] BakeDesc (slot, presence mask)
    int bake;                    // compose kernel: items are pose slots, output goes to c_rgb8 / c_px_all
    uint8_t *c_gsig;             // [n_slots][H*V/4] per 4-pixel group: signature of the foods that can tint it
    uint8_t *c_colhits;          // [n_slots][H]     transparent crossings recorded for the column
    void *c_hits;                // [n_slots][H][max_hits] HitRec
    void *dyn;                   // [n] EnvDyn, written by the logic kernel, read by the compose kernel
    void *hit_scratch;           // [grid][H][max_hits] HitRec when c.hits_in_global
    const int32_t *act;
    const float *act_c;          // continuous maze: [n][2] (turn_rate, walk_speed)
    float2 *cpos;                // continuous maze: position (float32, dynamics.py:75)
    double *cori;                // continuous maze: heading (python float)
    const double *coltab_d;      // [2][res_h]: cos_hp, sin_hp in float64 (heading independent)
    void *obs;
    double *rew;
    uint8_t *done;
    const uint8_t *mask;
    int do_step;                 // 0: observe only (reset), 1: step then observe
    int do_parts;                // compose kernel: image slices per env
    int T;                       // maze2d rollout: steps per launch
    uint64_t act_seed;
    uint32_t t_base;
    int32_t *act_out;
    MgbMirrors mir;              // maze2d rollout: every output is also stored at ptr + mir.delta[i]
    int auto_reset;
};

struct Env {
    int gx, gy, ori, steps;
    double life;
};

__device__ __forceinline__ const TaskHdr *blob_hdr(const uint8_t *b) { return reinterpret_cast<const TaskHdr *>(b); }

__device__ __forceinline__ void maze_evaluate(const MazeConst &c, const uint8_t *blob, int32_t *eaten, int64_t estride,
                                              Env &e, double &reward, int &done);

// action + evaluation_rule for one env (single thread).  `eaten` is strided by `estride` (SoA in HBM).
__device__ __forceinline__ void maze_logic(const MazeConst &c, const uint8_t *blob, int32_t *eaten, int64_t estride,
                                           Env &e, int action, double &reward, int &done)
{
    const TaskHdr *th = blob_hdr(blob);
    const int8_t *walls = reinterpret_cast<const int8_t *>(blob + c.off_walls);
    const int n = c.n;
    action &= 3;
    if (c.kind == MGB_MAZE_2D) {                                  // maze_2d.py:21-34 with DISCRETE_ACTIONS (dx, dy)
        int tx = e.gx + (action == 0 ? -1 : (action == 1 ? 1 : 0));
        int ty = e.gy + (action == 2 ? -1 : (action == 3 ? 1 : 0));
        if (tx < 0) tx += n;                                      // numpy negative index; unreachable (border walls)
        if (ty < 0) ty += n;
        if (tx < n && ty < n && walls[tx * n + ty] < 1) { e.gx = tx; e.gy = ty; }
    } else {                                                      // maze_discrete_3d.py:51-81, (turn, move)
        const int turn = action == 0 ? -1 : (action == 1 ? 1 : 0);
        const int mv = action == 2 ? -1 : (action == 3 ? 1 : 0);
        e.ori = (e.ori + turn + 4) & 3;
        int tx = e.gx, ty = e.gy;
        if (e.ori == 0) tx += mv; else if (e.ori == 1) ty += mv; else if (e.ori == 2) tx -= mv; else ty -= mv;
        if (tx >= 0 && tx < n && ty >= 0 && ty < n && walls[tx * n + ty] == 0) { e.gx = tx; e.gy = ty; }
    }
    maze_evaluate(c, blob, eaten, estride, e, reward, done);
}

// MazeBase.evaluation_rule (maze_base.py:65-95) on the agent's current cell
__device__ __forceinline__ void maze_evaluate(const MazeConst &c, const uint8_t *blob, int32_t *eaten, int64_t estride,
                                              Env &e, double &reward, int &done)
{
    const TaskHdr *th = blob_hdr(blob);
    const int n = c.n;
    e.steps += 1;                                                 // maze_base.py:66
    const bool over = e.steps > c.max_steps - 1;                  // :191-192
    if (c.task_type == MGB_MAZE_SURVIVAL) {
        double r = 0.0;
        const int8_t *fidx = reinterpret_cast<const int8_t *>(blob + c.off_fidx);
        const int f = fidx[e.gx * n + e.gy];
        if (f >= 0) {
            const double val = reinterpret_cast<const double *>(blob + c.off_fval)[f];
            const int itv = reinterpret_cast<const int32_t *>(blob + c.off_fint)[f];
            const int ea = eaten[f * estride];
            const bool present = (ea == kNever) || (e.steps > ea + itv);
            if (present && val > 1.0e-2) {                        // :71-75
                r = val;
                eaten[f * estride] = e.steps;
            }
        }
        e.life = e.life + (r + th->step_reward);                  // :78
        e.life = e.life < th->max_life ? e.life : th->max_life;   // :79
        done = (e.life < 0.0) || over;                            // :80
        reward = r;
    } else {
        const int goal = (e.gx == th->goal[0] && e.gy == th->goal[1]);
        reward = th->step_reward + (double)goal * th->goal_reward;   // :91-92
        done = goal || over;
    }
}

__device__ __forceinline__ void env_reset(const MazeConst &c, const uint8_t *blob, int32_t *eaten, int64_t estride,
                                          Env &e)
{
    const TaskHdr *th = blob_hdr(blob);
    e.gx = th->start[0]; e.gy = th->start[1]; e.ori = 0; e.steps = 0;
    e.life = th->initial_life;
    for (int f = 0; f < c.f_max; ++f) eaten[f * estride] = kNever;
}

// current transparent value of cell (i, j): SURVIVAL = remaining food (alias at maze_base.py:57), ESCAPE = goal one-hot
__device__ __forceinline__ double food_now(const MazeConst &c, const uint8_t *blob, const int32_t *eaten,
                                           int64_t estride, int steps, int cell)
{
    const int8_t *fidx = reinterpret_cast<const int8_t *>(blob + c.off_fidx);
    const int f = fidx[cell];
    if (f < 0) return 0.0;
    const int itv = reinterpret_cast<const int32_t *>(blob + c.off_fint)[f];
    const int ea = eaten[f * estride];
    const bool visible = (ea == kNever) || (steps >= ea + itv);
    return visible ? reinterpret_cast<const double *>(blob + c.off_fval)[f] : 0.0;
}


// ---------------------------------------------------------------------------------------------------------------
// MetaMazeContinuous3D dynamics (dynamics.py:16-92), float32/float64 typing of the numba + numpy original for float32
// actions: see oracle/maze_oracle.c (mo_vector_move_with_collision) for the line-by-line restatement this mirrors.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum2f(float a, float b) { float c = 0.0f; c += a; c += b; return c; }

__device__ float nearest_point_dev(const float pos[2], float l1x, float l1y, float l2x, float l2y, float np_out[2])
{
    float ux = l2x - l1x, uy = l2y - l1y;
    const float edge_norm = sqrtf(sum2f(ux * ux, uy * uy));
    const double den = 1.0e-6 > (double)edge_norm ? 1.0e-6 : (double)edge_norm;
    ux = (float)((double)ux / den); uy = (float)((double)uy / den);
    const float dist_1 = sum2f((pos[0] - l1x) * ux, (pos[1] - l1y) * uy);
    float qx, qy;
    if (dist_1 > edge_norm) { qx = l2x; qy = l2y; }
    else if (dist_1 < 0) { qx = l1x; qy = l1y; }
    else { qx = l1x + dist_1 * ux; qy = l1y + dist_1 * uy; }
    const float a = pos[0] - qx, b = pos[1] - qy;
    np_out[0] = qx; np_out[1] = qy;
    return sqrtf(sum2f(a * a, b * b));
}

__device__ void collision_force_dev(const float dv[2], double cell_size, double col_dist, float out[2])
{
    double dist = (double)sqrtf(sum2f(dv[0] * dv[0], dv[1] * dv[1]));
    const double eff = col_dist / cell_size;
    out[0] = out[1] = 0.0f;
    if (dist > 0.708 + eff) return;
    if (fabsf(dv[0]) < 0.5f && fabsf(dv[1]) < 0.5f) {
        const float k = (float)(0.50 / (dist > 1.0e-6 ? dist : 1.0e-6) * (0.708 + eff - dist) * cell_size);
        out[0] = k * dv[0]; out[1] = k * dv[1];
        return;
    }
    const bool x_pos = (dv[0] + dv[1] > 0), y_pos = (dv[1] - dv[0] > 0);
    float np[2];
    if (x_pos && y_pos) dist = (double)nearest_point_dev(dv, 0.5f, 0.5f, -0.5f, 0.5f, np);
    else if (!x_pos && y_pos) dist = (double)nearest_point_dev(dv, -0.5f, 0.5f, -0.5f, -0.5f, np);
    else if (!x_pos && !y_pos) dist = (double)nearest_point_dev(dv, -0.5f, -0.5f, 0.5f, -0.5f, np);
    else dist = (double)nearest_point_dev(dv, 0.5f, -0.5f, 0.5f, 0.5f, np);
    if (eff < dist) return;
    float ox = dv[0] - np[0], oy = dv[1] - np[1];
    const float on = sqrtf(sum2f(ox * ox, oy * oy));
    const double inv = 1.0 / (1.0e-6 > (double)on ? 1.0e-6 : (double)on);
    ox = (float)((double)ox * inv); oy = (float)((double)oy * inv);
    const float k = (float)(0.50 * (eff - dist) * cell_size);
    out[0] = k * ox; out[1] = k * oy;
}

// do_action (maze_continuous_3d.py:47-53): clip, scale, ten sub-steps of vector_move + collision forces
__device__ void continuous_move(const MazeConst &c, const uint8_t *blob, float tr, float ws, float pos[2], double &ori_io)
{
    const TaskHdr *th = blob_hdr(blob);
    const int8_t *walls = reinterpret_cast<const int8_t *>(blob + c.off_walls);
    const int n = c.n;
    const double cell_size = th->cell_size;
    float turn = tr < -1.0f ? -1.0f : (tr > 1.0f ? 1.0f : tr);
    turn = turn * (float)3.1415926;
    float walk = ws < -1.0f ? -1.0f : (ws > 1.0f ? 1.0f : ws);
    if (walk < 0) walk = walk * 0.5f;
    float tx = pos[0], ty = pos[1];
    double ori = ori_io;
    for (int it = 0; it < 10; ++it) {                       // int(100 * 0.10), dynamics.py:76
        double fin = ori + (double)turn * 0.01;
        const double off_ori = 0.5 * (fin + ori);
        const double off = (double)walk * 0.01;
        const float dx = (float)(cos(off_ori) * off), dy = (float)(sin(off_ori) * off);
        while (fin > 6.2831852) fin -= 6.2831852;
        while (fin < 0) fin += 6.2831852;
        ori = fin;
        const float ex = tx + dx, ey = ty + dy;
        const float cs = (float)cell_size;
        const float ecx = ex / cs, ecy = ey / cs;
        float cx = 0.0f, cy = 0.0f;
        for (int i = -1; i < 2; ++i)
            for (int j = -1; j < 2; ++j) {
                const int wi = i + (int)ecx, wj = j + (int)ecy;
                if (wi > -1 && wi < n && wj > -1 && wj < n && walls[wi * n + wj] > 0) {
                    const float dv[2] = {ecx - floorf(ecx) - (float)(i + 0.5), ecy - floorf(ecy) - (float)(j + 0.5)};
                    float f[2];
                    collision_force_dev(dv, cell_size, 0.20, f);   // collision_dist = 0.20, maze_continuous_3d.py:20
                    cx += f[0]; cy += f[1];
                }
            }
        tx = cx + ex; ty = cy + ey;
    }
    pos[0] = tx; pos[1] = ty;
    ori_io = ori;
}

// ---------------------------------------------------------------------------------------------------------------
// state-only kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void maze_reset_kernel(const __grid_constant__ MazeConst c, const __grid_constant__ MazeArgs a)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    if (a.mask && !a.mask[e]) return;
    const uint8_t *blob = a.blobs + (int64_t)a.env2task[e] * c.blob_bytes;
    Env s;
    env_reset(c, blob, a.eaten + e, a.n_pad, s);
    a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
    a.life[e] = s.life;
    if (c.kind == MGB_MAZE_CONTINUOUS_3D) {             // get_cell_center(start), heading 0 (maze_base.py:41,50)
        const TaskHdr *th = blob_hdr(blob);
        a.cpos[e] = make_float2((float)(s.gx * th->cell_size + 0.5 * th->cell_size),
                                (float)(s.gy * th->cell_size + 0.5 * th->cell_size));
        a.cori[e] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2-D: one thread per env; observation tile of the CTA leaves through shared memory + one bulk store
// ---------------------------------------------------------------------------------------------------------------
constexpr int k2dThreads = 128;

__global__ void __launch_bounds__(k2dThreads) maze2d_kernel(const __grid_constant__ MazeConst c,
                                                            const __grid_constant__ MazeArgs a)
{
    extern __shared__ __align__(128) float tile2d[];
    const int64_t e0 = (int64_t)blockIdx.x * k2dThreads;
    const int64_t e = e0 + threadIdx.x;
    const int W = 2 * c.view_grid + 1, D = W * W;
    const int rows = (int)((a.n - e0) < k2dThreads ? (a.n - e0) : k2dThreads);
    if (e < a.n) {
        const uint8_t *blob = a.blobs + (int64_t)a.env2task[e] * c.blob_bytes;
        const TaskHdr *th = blob_hdr(blob);
        const int4 ag = a.agent[e];
        Env s = {ag.x, ag.y, ag.z, ag.w, a.life[e]};
        int32_t *eaten = a.eaten + e;
        if (a.do_step) {
            double reward;
            int done;
            maze_logic(c, blob, eaten, a.n_pad, s, a.act[e], reward, done);
            a.rew[e] = reward;
            a.done[e] = (uint8_t)done;
            if (done && a.auto_reset) env_reset(c, blob, eaten, a.n_pad, s);
            a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
            a.life[e] = s.life;
        }
        // update_observation, maze_2d.py:89-121
        const int8_t *walls = reinterpret_cast<const int8_t *>(blob + c.off_walls);
        const int n = c.n, g = c.view_grid;
        float *row = tile2d + threadIdx.x * D;
        for (int p = 0; p < W; ++p)
            for (int q = 0; q < W; ++q) {
                const int x = s.gx - g + p, y = s.gy - g + q;
                float v = -1.0f;
                if (x >= 0 && x < n && y >= 0 && y < n) {
                    v = (float)(-(int)walls[x * n + y]);
                    if (c.task_type == MGB_MAZE_SURVIVAL)
                        v = (float)((double)v + food_now(c, blob, eaten, a.n_pad, s.steps, x * n + y));
                    else
                        v = (float)((double)v + ((x == th->goal[0] && y == th->goal[1]) ? 1.0 : 0.0));
                }
                row[p * W + q] = v;
            }
        if (c.task_type == MGB_MAZE_SURVIVAL) row[g * W + g] = (float)s.life;
    }
    float *dst = reinterpret_cast<float *>(a.obs) + e0 * D;
    const uint32_t bytes = (uint32_t)rows * (uint32_t)D * 4u;
    if ((bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
        mgb_fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            mgb_bulk_store(dst, tile2d, bytes);
            mgb_bulk_commit();
            mgb_bulk_wait_read<0>();   // smem must outlive the copy; the kernel boundary flushes the writes
        }
    } else {
        __syncthreads();
        for (int i = threadIdx.x; i < rows * D; i += blockDim.x) dst[i] = tile2d[i];
    }
}


// T MetaMaze2D steps in one launch: the agent (cell, step counter, life) stays in registers, food stamps stay in their
// SoA slots, each step's observation tile of the CTA leaves through double-buffered shared memory + one bulk store.
template <int XM>   // 0 plain, 1 peer mirrors, 2 multicast-only stores (see quad_rollout_kernel)
__global__ void __launch_bounds__(k2dThreads) maze2d_rollout_kernel(const __grid_constant__ MazeConst c,
                                                                    const __grid_constant__ MazeArgs a)
{
    extern __shared__ __align__(128) float tile2d[];
    const int64_t e0 = (int64_t)blockIdx.x * k2dThreads;
    const int64_t e = e0 + threadIdx.x;
    const int W = 2 * c.view_grid + 1, D = W * W;
    const int rows = (int)((a.n - e0) < k2dThreads ? (a.n - e0) : k2dThreads);
    const bool active = e < a.n;
    const uint8_t *blob = nullptr;
    const TaskHdr *th = nullptr;
    const int8_t *walls = nullptr;
    Env s = {0, 0, 0, 0, 0.0};
    int32_t *eaten = a.eaten + e;
    if (active) {
        blob = a.blobs + (int64_t)a.env2task[e] * c.blob_bytes;
        th = blob_hdr(blob);
        walls = reinterpret_cast<const int8_t *>(blob + c.off_walls);
        const int4 ag = a.agent[e];
        s.gx = ag.x; s.gy = ag.y; s.ori = ag.z; s.steps = ag.w; s.life = a.life[e];
    }
    const uint2 akey = make_uint2((uint32_t)a.act_seed, (uint32_t)(a.act_seed >> 32));
    const int64_t genv = a.env_base + e;
    const int n = c.n, g = c.view_grid;
    for (int t = 0; t < a.T; ++t) {
        float *tile = tile2d + (size_t)(t & 1) * k2dThreads * D;
        if (threadIdx.x == 0) mgb_bulk_wait_read<1>();      // the store issued two steps ago has read this tile
        __syncthreads();
        uint32_t done_byte = 0;
        if (active) {
            int action;
            if (a.act) action = a.act[(int64_t)t * a.n + e];
            else {
                const uint4 r = mgb_philox4x32_10(make_uint4((uint32_t)genv, (uint32_t)((uint64_t)genv >> 32),
                                                             a.t_base + (uint32_t)t, MGB_STREAM_ACTION), akey);
                action = (int)(r.x >> 30);                   // uniform over {0, 1, 2, 3}
                if (a.act_out) {
                    if (XM == 2) mgb_mc_st(mgb_shift(a.act_out + (int64_t)t * a.n + e, a.mir.delta[0]), (int32_t)action);
                    else a.act_out[(int64_t)t * a.n + e] = action;
                    if (XM == 1) mgb_mirror_store(a.mir, a.act_out + (int64_t)t * a.n + e, (int32_t)action);
                }
            }
            double reward;
            int done;
            maze_logic(c, blob, eaten, a.n_pad, s, action, reward, done);
            if (done && a.auto_reset) env_reset(c, blob, eaten, a.n_pad, s);
            if (a.rew) {
                if (XM == 2) mgb_mc_st(mgb_shift(a.rew + (int64_t)t * a.n + e, a.mir.delta[0]), reward);
                else a.rew[(int64_t)t * a.n + e] = reward;
                if (XM == 1) mgb_mirror_store(a.mir, a.rew + (int64_t)t * a.n + e, reward);
            }
            done_byte = (uint32_t)done;
            if (a.done && XM != 2) {
                a.done[(int64_t)t * a.n + e] = (uint8_t)done;
                if (XM == 1) mgb_mirror_store(a.mir, a.done + (int64_t)t * a.n + e, (uint8_t)done);
            }
            if (a.obs) {
                float *row = tile + threadIdx.x * D;
                for (int p = 0; p < W; ++p)
                    for (int q = 0; q < W; ++q) {
                        const int x = s.gx - g + p, y = s.gy - g + q;
                        float v = -1.0f;
                        if (x >= 0 && x < n && y >= 0 && y < n) {
                            v = (float)(-(int)walls[x * n + y]);
                            if (c.task_type == MGB_MAZE_SURVIVAL)
                                v = (float)((double)v + food_now(c, blob, eaten, a.n_pad, s.steps, x * n + y));
                            else
                                v = (float)((double)v + ((x == th->goal[0] && y == th->goal[1]) ? 1.0 : 0.0));
                        }
                        row[p * W + q] = v;
                    }
                if (c.task_type == MGB_MAZE_SURVIVAL) row[g * W + g] = (float)s.life;
            }
        }
        if (XM == 2) {
            if (a.done) mgb_mc_st_bytes(mgb_shift(a.done + (int64_t)t * a.n + e, a.mir.delta[0]), done_byte, active);
            if (a.obs) {
                __syncthreads();
                mgb_mc_copy_tile(mgb_shift(reinterpret_cast<float *>(a.obs) + ((int64_t)t * a.n + e0) * D, a.mir.delta[0]),
                                 tile, (uint32_t)rows * (uint32_t)D * 4u);
            }
        } else if (a.obs) {
            float *dst = reinterpret_cast<float *>(a.obs) + ((int64_t)t * a.n + e0) * D;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)D * 4u;
            if ((bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
                mgb_fence_proxy_async();
                __syncthreads();
                if (threadIdx.x == 0) {
                    mgb_bulk_store(dst, tile, bytes);
                    if (XM == 1) mgb_mirror_bulk_store(a.mir, dst, tile, bytes);
                    mgb_bulk_commit();
                }
            } else {
                __syncthreads();
                for (int i = threadIdx.x; i < rows * D; i += blockDim.x) {
                    dst[i] = tile[i];
                    if (XM == 1) mgb_mirror_store(a.mir, dst + i, tile[i]);
                }
            }
        }
    }
    if (active) {
        a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
        a.life[e] = s.life;
    }
    if (threadIdx.x == 0) mgb_bulk_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// discrete 3-D: persistent CTA, step logic + raycast render
// ---------------------------------------------------------------------------------------------------------------
struct ColRec {                  // one screen column (ray), 64 bytes
    double cos_hp, cos_abs, sin_abs;   // float32 table values promoted (ray_caster_utils.py:82-92)
    double light, oma, ratio;          // wall shading: |cos/sin|, 1 - alpha, hit_dist * cos_hp / l_focal
    int16_t v_s, v_e;                  // wall span [v_s, v_e)
    int16_t ti, text_id;               // texture column, texture id
    int16_t n_hits, wall;              // transparent crossings, wall hit within max_vision
    int32_t pad;
};
struct RowRec {                  // one screen row, 24 bytes
    double distance, light;
    int32_t kind;                // 0 none, 1 floor, 2 ceiling
    int32_t pad;
};
struct HitRec {                  // one transparent crossing of a column, 16 bytes
    double tf;
    int16_t v_s, v_e;
    int32_t fid;                 // food slot of the crossed cell (pose cache: presence is decided per env, per step)
};
struct EnvDyn {                  // per env, per step: what the static pose layers must be combined with, 40 bytes
    int32_t slot;                // pose-cache slot of (task, cell, heading)
    int32_t bar_end;             // life bar end column (python slice semantics already applied)
    uint64_t present[2];         // bit f: food slot f is currently visible
    int32_t task, pad;           // pad: 8-bit signature of the missing foods that can tint this pose (0 = frame is final)
    int32_t vframe, pad2;        // >= 0: finished frame c_var8[vframe]; -1: c_rgb8[slot] (+ the tints `pad` asks for)
};
// what make_dyn needs to know about a pose, in ONE 32-byte record indexed like pose_index (two 16-byte loads issued together
// instead of the dependent chain pose_index -> c_fmask[slot] -> c_vbase[slot])
struct PoseRec {
    int32_t slot;                // pose-cache slot, -1: wall cell (never an agent pose)
    int32_t vbase;               // first variant frame of the pose, -1: none
    uint64_t fmask[2];           // food slots that can change this pose's image
    uint64_t pad;
};
struct BakeDesc {                // one variant frame to bake
    int32_t slot, pad;
    uint64_t present[2];
};
constexpr int kVariantBitsMax = 7;

// finished uint8 frame the env's observation starts from
__device__ __forceinline__ const uint8_t *frame_of(const MazeArgs &a, const EnvDyn &d, size_t frame_bytes)
{
    return d.vframe >= 0 ? a.c_var8 + (size_t)d.vframe * frame_bytes : a.c_rgb8 + (size_t)d.slot * frame_bytes;
}

__device__ __forceinline__ int trunc_i(double x) { return (int)x; }   // cvt.rzi: python/numba int()
// 32-bit observation word of a pixel value: the integer itself (MGB_OBS_I32, what the reference's array holds) or the same
// value as float32 (MGB_OBS_F32, the dtype the reference's observation_space declares, maze_env.py:37-39)
#define MGB_OBS_WORD(c, v) ((c).obs_dtype == MGB_OBS_F32 ? __float_as_int((float)(v)) : (v))

// rgb = light * (alpha * FAR_RGB + (1 - alpha) * texel), FAR_RGB = 0 (ray_caster_utils.py:7,118)
__device__ __forceinline__ void shade(int rgb[3], double light, double oma, uint32_t texel)
{
    rgb[0] = trunc_i(light * (oma * (double)(texel & 0xffu)));
    rgb[1] = trunc_i(light * (oma * (double)((texel >> 8) & 0xffu)));
    rgb[2] = trunc_i(light * (oma * (double)((texel >> 16) & 0xffu)));
}
// rgb = (1 - tf) * rgb + tf * (0, 255, 0)  (ray_caster_utils.py:8,121-123)
__device__ __forceinline__ void blend(int rgb[3], double tf)
{
    const double k = 1.0 - tf;
    rgb[0] = trunc_i(k * (double)rgb[0] + tf * 0.0);
    rgb[1] = trunc_i(k * (double)rgb[1] + tf * 255.0);
    rgb[2] = trunc_i(k * (double)rgb[2] + tf * 0.0);
}

__device__ __forceinline__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// FILL = false: render the observation of every env directly (step logic + ray cast + paint).
// FILL = true : render the STATIC layers of every cached pose (task, cell, heading) once, at set_task time: colours
//               before any transparency, the food slot under every floor/ceiling pixel, every POSSIBLE transparent
//               crossing of every column.  maze3d_compose_kernel then turns a pose + the env's current food state
//               into the exact observation with integer work only (the float64 geometry is memoised).
template <bool FILL>
__global__ void __launch_bounds__(kRenderThreads, 1) maze3d_kernel(const __grid_constant__ MazeConst c,
                                                                   const __grid_constant__ MazeArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int n = c.n, H = c.res_h, V = c.res_v, ts = c.ts;
    const int tex_words = (c.n_tex + 1) * ts * ts;
    // ---- shared memory carve-up (mirrors maze3d_smem_bytes on the host)
    size_t off = 0;
    uint32_t *s_tex = reinterpret_cast<uint32_t *>(smem + off);          off = align_up(off + (size_t)tex_words * 4, 128);
    // record sets: [0] always, [1] when the direct renderer pipelines geometry of env e + 1 under the pixels of env e (c.pipe)
    const int nbuf = (!FILL && c.pipe) ? 2 : 1;
    uint8_t *s_blob2[2];
    double *s_transp2[2];
    ColRec *s_col2[2];
    RowRec *s_row2[2];
    HitRec *s_hit2[2];
    for (int b = 0; b < 2; ++b) {                                         // maze tiles: always two (this env's + the next one's in flight)
        s_blob2[b] = smem + off;                                          off = align_up(off + c.blob_bytes, 128);
    }
    for (int b = 0; b < 2; ++b) {
        if (b < nbuf) {
            s_transp2[b] = reinterpret_cast<double *>(smem + off);       off = align_up(off + (size_t)n * n * 8, 128);
            s_col2[b] = reinterpret_cast<ColRec *>(smem + off);          off = align_up(off + (size_t)H * sizeof(ColRec), 128);
            s_row2[b] = reinterpret_cast<RowRec *>(smem + off);          off = align_up(off + (size_t)V * sizeof(RowRec), 128);
            if (c.hits_in_global) s_hit2[b] = reinterpret_cast<HitRec *>(a.hit_scratch) + ((size_t)blockIdx.x * nbuf + b) * H * c.max_hits;
            else { s_hit2[b] = reinterpret_cast<HitRec *>(smem + off); off = align_up(off + (size_t)H * c.max_hits * sizeof(HitRec), 128); }
        } else {
            s_transp2[b] = s_transp2[0]; s_col2[b] = s_col2[0]; s_row2[b] = s_row2[0]; s_hit2[b] = s_hit2[0];
        }
    }
    double *s_rowc = reinterpret_cast<double *>(smem + off);             off = align_up(off + (size_t)V * 8, 128);
    const int px_bytes = c.obs_dtype == MGB_OBS_U8 ? 3 : 12;
    const int run_bytes = c.run_px * px_bytes;                           // 768
    const int n_slots = c.obs_dtype == MGB_OBS_U8 ? 2 : 1;               // int32 runs are 4x larger: single slot
    uint8_t *s_out = smem + off;                                          off = align_up(off + (size_t)(kRenderThreads / 32) * n_slots * run_bytes, 128);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + off);          off += 32;
    int *s_env2[2];                                                       // [0..3] gx gy ori steps, [4] lifebar end
    double *s_pose2[2];                                                   // continuous maze: x, y, sin(ori), cos(ori)
    s_env2[0] = reinterpret_cast<int *>(smem + off);      s_pose2[0] = reinterpret_cast<double *>(smem + off + 32);
    s_env2[1] = reinterpret_cast<int *>(smem + off + 64); s_pose2[1] = reinterpret_cast<double *>(smem + off + 96);
    int *s_runctr = reinterpret_cast<int *>(smem + off + 128);            // [b]: next column run of record set b (pipelined mode)

    const int tid = threadIdx.x;
    // screen-row centre above the horizon, half_v - (d_v + 0.5) * pixel_size: the wall texel's row term (:184), pose independent
    for (int d_v = tid; d_v < V; d_v += blockDim.x) s_rowc[d_v] = c.half_v - (d_v + 0.5) * c.pixel_size;
    if (tid == 0) {
        mgb_mbar_init(&s_bar[0], 1);   // textures
        mgb_mbar_init(&s_bar[1], 1);   // task blob, buffer 0
        mgb_mbar_init(&s_bar[2], 1);   // task blob, buffer 1
        mgb_fence_mbar_init();
    }
    __syncthreads();
    // ---- textures: ONE bulk (TMA) copy per CTA for the whole launch, in <= 64 KB pieces
    if (tid == 0) {
        const uint32_t total = (uint32_t)tex_words * 4u;
        mgb_mbar_expect_tx(&s_bar[0], total);
        for (uint32_t o = 0; o < total; o += 65536u) {
            const uint32_t len = total - o < 65536u ? total - o : 65536u;
            mgb_bulk_load(reinterpret_cast<uint8_t *>(s_tex) + o, reinterpret_cast<const uint8_t *>(a.tex) + o, len,
                          &s_bar[0]);
        }
    }
    uint32_t blob_phase = 0;      // bit b: parity of buffer b's mbarrier
    bool tex_ready = false;
    int run_parity = 0;

    // ---- maze tiles (walls, texture ids, food table) arrive by TMA one env ahead of their use
    auto load_blob = [&](int64_t e, int b) {
        const int task_id = FILL ? a.poses[e].x : a.env2task[e];
        const uint8_t *src = a.blobs + (int64_t)task_id * c.blob_bytes;
        mgb_mbar_expect_tx(&s_bar[1 + b], (uint32_t)c.blob_bytes);
        mgb_bulk_load(b ? s_blob2[1] : s_blob2[0], src, (uint32_t)c.blob_bytes, &s_bar[1 + b]);
    };

    // ---- per-env geometry: tile, pose, transparent map, row table, one DDA ray per column -> record set `b`.  Run by `gn`
    // threads (gt = index among them) that synchronise among themselves: the whole CTA (__syncthreads), or -- pipelined
    // direct renderer -- the first kGeoThreads threads (named barrier 2) while every warp paints the previous env.
    // Tile buffer bb: loaded here (pipelined mode: the other buffer is being read by the pixel warps), or already requested by
    // the previous call, which prefetches `next_e` into the other buffer after its own wait (sequential mode).
    auto geometry = [&](int64_t e, int b, int bb, int gt, int gn, bool named, bool load_here, int64_t next_e) {
        auto gsync = [&]() { if (named) asm volatile("bar.sync 2, %0;" ::"n"(kGeoThreads) : "memory"); else __syncthreads(); };
        uint8_t *s_blob = (bb ? s_blob2[1] : s_blob2[0]);
        double *s_transp = (b ? s_transp2[1] : s_transp2[0]);
        ColRec *s_col = (b ? s_col2[1] : s_col2[0]);
        RowRec *s_row = (b ? s_row2[1] : s_row2[0]);
        HitRec *s_hit = (b ? s_hit2[1] : s_hit2[0]);
        int *s_env = (b ? s_env2[1] : s_env2[0]);
        double *s_pose = (b ? s_pose2[1] : s_pose2[0]);
        if (gt == 0) { if (load_here) load_blob(e, bb); s_runctr[b] = 0; }
        mgb_mbar_wait(&s_bar[1 + bb], (blob_phase >> bb) & 1u);
        blob_phase ^= 1u << bb;
        if (gt == 0 && next_e >= 0) load_blob(next_e, bb ^ 1);
        const TaskHdr *th = blob_hdr(s_blob);
        // ---- step logic (one thread), then publish agent pose to the CTA
        if (FILL) {
            if (gt == 0) {
                const int4 ps = a.poses[e];
                s_env[0] = ps.y; s_env[1] = ps.z; s_env[2] = ps.w; s_env[3] = 0; s_env[4] = 0;
            }
        } else if (gt == 0) {
            const int4 ag = a.agent[e];
            Env s = {ag.x, ag.y, ag.z, ag.w, a.life[e]};
            int32_t *eaten = a.eaten + e;
            const bool cont = c.kind == MGB_MAZE_CONTINUOUS_3D;
            float cp[2] = {0.f, 0.f};
            double co = 0.0;
            if (cont) { const float2 p2 = a.cpos[e]; cp[0] = p2.x; cp[1] = p2.y; co = a.cori[e]; }
            if (a.do_step) {
                double reward;
                int done;
                if (cont) {
                    continuous_move(c, s_blob, a.act_c[2 * e], a.act_c[2 * e + 1], cp, co);
                    const float csf = (float)th->cell_size;                 // get_loc_grid, maze_base.py:199-202
                    s.gx = (int)(cp[0] / csf); s.gy = (int)(cp[1] / csf);
                    maze_evaluate(c, s_blob, eaten, a.n_pad, s, reward, done);
                } else {
                    maze_logic(c, s_blob, eaten, a.n_pad, s, a.act[e], reward, done);
                }
                a.rew[e] = reward;
                a.done[e] = (uint8_t)done;
                if (done && a.auto_reset) {
                    env_reset(c, s_blob, eaten, a.n_pad, s);
                    if (cont) {
                        cp[0] = (float)(s.gx * th->cell_size + 0.5 * th->cell_size);
                        cp[1] = (float)(s.gy * th->cell_size + 0.5 * th->cell_size);
                        co = 0.0;
                    }
                }
                a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
                a.life[e] = s.life;
                if (cont) { a.cpos[e] = make_float2(cp[0], cp[1]); a.cori[e] = co; }
            }
            if (cont) {          // publish the pose: position promoted from float32, sin/cos of the float64 heading
                s_pose[0] = (double)cp[0]; s_pose[1] = (double)cp[1]; s_pose[2] = sin(co); s_pose[3] = cos(co);
            }
            s_env[0] = s.gx; s_env[1] = s.gy; s_env[2] = s.ori; s_env[3] = s.steps;
            // life bar extent, maze_discrete_3d.py:118-126 (python slice semantics on the end index)
            int ex = trunc_i(c.lb_sx + s.life / th->max_life * c.lb_l);
            if (ex < 0) { ex += H; if (ex < 0) ex = 0; }
            if (ex > H) ex = H;
            s_env[4] = ex;
        }
        gsync();
        const int gx = s_env[0], gy = s_env[1], ori = s_env[2], steps = s_env[3];
        const double cell_size = th->cell_size, vision_height = th->agent_height, ceil_height = th->wall_height;
        const bool cont_pose = !FILL && c.kind == MGB_MAZE_CONTINUOUS_3D;
        const double pos_x = cont_pose ? s_pose[0] : gx * cell_size + 0.5 * cell_size;   // get_cell_center, maze_base.py:194-197
        const double pos_y = cont_pose ? s_pose[1] : gy * cell_size + 0.5 * cell_size;
        const double text_to_cell = c.text_size / cell_size;
        (void)text_to_cell;

        // ---- transparent map + row table
        for (int k = gt; k < n * n; k += gn) {
            double v;
            if (c.task_type == MGB_MAZE_SURVIVAL) {
                if (FILL) {   // static superset: every food cell at its full value
                    const int f = reinterpret_cast<const int8_t *>(s_blob + c.off_fidx)[k];
                    v = f >= 0 ? reinterpret_cast<const double *>(s_blob + c.off_fval)[f] : 0.0;
                } else v = food_now(c, s_blob, a.eaten + e, a.n_pad, steps, k);
            } else v = (k == th->goal[0] * n + th->goal[1]) ? 1.0 : 0.0;
            s_transp[k] = v;
        }
        for (int d_v = gt; d_v < V; d_v += gn) {
            RowRec r;
            r.kind = 0; r.distance = 0.0; r.light = 0.0; r.pad = 0;
            if (d_v > V / 2) {                                       // floor rows, ray_caster_utils.py:95-101
                const double v_screen = (d_v + 0.5) * c.pixel_size - c.half_v;
                r.distance = vision_height / v_screen * c.l_focal;
                r.light = v_screen / c.l_focal;
                r.kind = r.distance > c.max_vision ? 0 : 1;
            } else if (d_v < V / 2) {                                // ceiling rows, :129-135
                const double v_screen = c.half_v - (d_v + 0.5) * c.pixel_size;
                r.distance = (ceil_height - vision_height) / v_screen * c.l_focal;
                r.light = v_screen / c.l_focal;
                r.kind = r.distance > c.max_vision ? 0 : 2;
            }
            s_row[d_v] = r;
        }
        gsync();
        // ---- one ray per column: DDA_2D (ray_caster_utils.py:11-62) + wall span set-up (:156-182)
        const int8_t *walls = reinterpret_cast<const int8_t *>(s_blob + c.off_walls);
        const int8_t *texts = reinterpret_cast<const int8_t *>(s_blob + c.off_texts);
        const float *ct = a.coltab + (size_t)(cont_pose ? 0 : ori) * 3 * H;
        for (int d_h = gt; d_h < H; d_h += gn) {
            ColRec cr;
            cr.cos_hp = (double)ct[d_h];
            if (cont_pose) {     // float64 heading: tables from the float64 sin/cos, stored as float32 (:82-92)
                const double ch = a.coltab_d[d_h], sh = a.coltab_d[H + d_h];
                cr.sin_abs = (double)(float)(sh * s_pose[3] + ch * s_pose[2]);
                cr.cos_abs = (double)(float)(ch * s_pose[3] - sh * s_pose[2]);
            } else {
                cr.cos_abs = (double)ct[H + d_h];
                cr.sin_abs = (double)ct[2 * H + d_h];
            }
            const double cos_ori = cr.cos_abs, sin_ori = cr.sin_abs;
            const int i0 = trunc_i(pos_x / cell_size), j0 = trunc_i(pos_y / cell_size);
            const double delta_dist_x = fabs(cos_ori) < 1.0e-6 ? 1.0e+6 : fabs(cell_size / cos_ori);
            const double delta_dist_y = fabs(sin_ori) < 1.0e-6 ? 1.0e+6 : fabs(cell_size / sin_ori);
            const double d_x = cos_ori > 0 ? ((i0 + 1) * cell_size - pos_x) : (i0 * cell_size - pos_x);
            const double d_y = sin_ori > 0 ? ((j0 + 1) * cell_size - pos_y) : (j0 * cell_size - pos_y);
            double side_x = fabs(cos_ori) < 1.0e-6 ? 1.0e+6 : d_x / cos_ori;
            double side_y = fabs(sin_ori) < 1.0e-6 ? 1.0e+6 : d_y / sin_ori;
            const int delta_i = cos_ori > 0 ? 1 : -1, delta_j = sin_ori > 0 ? 1 : -1;
            int hit_i = i0, hit_j = j0, hit_side = 0, nh = 0;
            double hit_dist = 0.0;
            HitRec *hits = s_hit + (size_t)d_h * c.max_hits;
            // a transparent crossing at distance d paints the span of a wall standing there (:191-198)
            const int8_t *fidx_map = reinterpret_cast<const int8_t *>(s_blob + c.off_fidx);
            auto record_hit = [&](double d, double strength, int cell) {
                if (nh >= c.max_hits) return;
                const double ratio = d * cr.cos_hp / c.l_focal;
                const double tv = (ceil_height - vision_height) / ratio, bv = vision_height / ratio;
                const int s0 = trunc_i((c.half_v - tv) / c.pixel_size), s1 = trunc_i((c.half_v + bv) / c.pixel_size);
                HitRec hr;
                hr.tf = strength * 0.50 + 0.10;
                hr.v_s = (int16_t)(s0 < 0 ? 0 : s0);
                hr.v_e = (int16_t)(s1 > V ? V : s1);
                hr.fid = c.task_type == MGB_MAZE_SURVIVAL ? (int)fidx_map[cell] : 0;
                hits[nh++] = hr;
            };
            if (s_transp[hit_i * n + hit_j] > 0.01)                   // start cell, :25-29
                record_hit(side_x < side_y ? side_x : side_y, s_transp[hit_i * n + hit_j], hit_i * n + hit_j);
            while (hit_dist < c.max_vision) {
                if (side_x < side_y) {
                    hit_i += delta_i;
                    side_y -= side_x;
                    hit_dist += side_x;
                    if (hit_i < 0 || hit_i >= n) {
                        if (hit_j < 0 || hit_j >= n) { hit_dist = 1.0e+6; break; }
                    } else if (hit_j >= 0 && hit_j < n) {
                        const double tv = s_transp[hit_i * n + hit_j];
                        if (tv > 0.01) record_hit(hit_dist, tv, hit_i * n + hit_j);
                        if (walls[hit_i * n + hit_j] > 0) { hit_side = 0; break; }
                    }
                    side_x = delta_dist_x;
                } else {
                    hit_j += delta_j;
                    side_x -= side_y;
                    hit_dist += side_y;
                    if (hit_i < 0 || hit_i >= n) {
                        if (hit_j < 0 || hit_j >= n) { hit_dist = 1.0e+6; break; }
                    } else if (hit_j >= 0 && hit_j < n) {
                        const double tv = s_transp[hit_i * n + hit_j];
                        if (tv > 0.01) record_hit(hit_dist, tv, hit_i * n + hit_j);
                        if (walls[hit_i * n + hit_j] > 0) { hit_side = 1; break; }
                    }
                    side_y = delta_dist_y;
                }
            }
            cr.wall = hit_dist > c.max_vision ? 0 : 1;               // :162-163 (skips overlays too)
            cr.n_hits = 0; cr.v_s = 0; cr.v_e = 0; cr.ti = 0; cr.text_id = 0;
            cr.light = 0.0; cr.oma = 0.0; cr.ratio = 1.0; cr.pad = 0;
            if (cr.wall) {
                const double alpha = fmin(1.0, fmax(2.0 * hit_dist / c.max_vision - 1.0, 0.0));
                const int ci = hit_i < 0 ? 0 : (hit_i >= n ? n - 1 : hit_i);
                const int cj = hit_j < 0 ? 0 : (hit_j >= n ? n - 1 : hit_j);
                cr.text_id = texts[ci * n + cj];
                const double hit_pt_x = hit_dist * cos_ori + pos_x;
                const double hit_pt_y = hit_dist * sin_ori + pos_y;
                double local_h;
                if (hit_side == 0) {
                    local_h = hit_pt_y / cell_size; local_h -= floor(local_h);
                    cr.light = fabs(cos_ori);
                } else {
                    local_h = hit_pt_x / cell_size; local_h -= floor(local_h);
                    cr.light = fabs(sin_ori);
                }
                double d_i = local_h / c.text_size;
                d_i -= floor(d_i);
                cr.ti = (int16_t)trunc_i(ts * d_i);
                cr.oma = 1.0 - alpha;
                cr.ratio = hit_dist * cr.cos_hp / c.l_focal;
                const double top_v = (ceil_height - vision_height) / cr.ratio, bot_v = vision_height / cr.ratio;
                int v_s = trunc_i((c.half_v - top_v) / c.pixel_size), v_e = trunc_i((c.half_v + bot_v) / c.pixel_size);
                cr.v_s = (int16_t)(v_s < 0 ? 0 : v_s);
                cr.v_e = (int16_t)(v_e > V ? V : v_e);
                cr.n_hits = (int16_t)nh;
            }
            s_col[d_h] = cr;
            if (FILL) {
                a.c_colhits[(size_t)e * H + d_h] = (uint8_t)cr.n_hits;
                HitRec *gh = reinterpret_cast<HitRec *>(a.c_hits) + ((size_t)e * H + d_h) * c.max_hits;
                for (int k = 0; k < cr.n_hits; ++k) gh[k] = hits[k];
            }
        }
    };
    // ---- pixels of env e from record set b (direct renderer).  dyn: warps pull column runs from a shared counter (the
    // geometry warps join late), else run = warp, warp + n_warps, ...
    auto pixels = [&](int64_t e, int b, int bb, bool dyn) {
        uint8_t *s_blob = (bb ? s_blob2[1] : s_blob2[0]);
        double *s_transp = (b ? s_transp2[1] : s_transp2[0]);
        ColRec *s_col = (b ? s_col2[1] : s_col2[0]);
        RowRec *s_row = (b ? s_row2[1] : s_row2[0]);
        HitRec *s_hit = (b ? s_hit2[1] : s_hit2[0]);
        int *s_env = (b ? s_env2[1] : s_env2[0]);
        double *s_pose = (b ? s_pose2[1] : s_pose2[0]);
        const TaskHdr *th = blob_hdr(s_blob);
        const int gx = s_env[0], gy = s_env[1], ori = s_env[2], steps = s_env[3];
        const double cell_size = th->cell_size, vision_height = th->agent_height, ceil_height = th->wall_height;
        const bool cont_pose = !FILL && c.kind == MGB_MAZE_CONTINUOUS_3D;
        const double pos_x = cont_pose ? s_pose[0] : gx * cell_size + 0.5 * cell_size;   // get_cell_center, maze_base.py:194-197
        const double pos_y = cont_pose ? s_pose[1] : gy * cell_size + 0.5 * cell_size;
        const double text_to_cell = c.text_size / cell_size;

        const int8_t *walls = reinterpret_cast<const int8_t *>(s_blob + c.off_walls);
        const int8_t *texts = reinterpret_cast<const int8_t *>(s_blob + c.off_texts);
        const float *ct = a.coltab + (size_t)(cont_pose ? 0 : ori) * 3 * H;
        (void)walls; (void)steps; (void)ct; (void)ceil_height; (void)text_to_cell; (void)s_transp; (void)s_hit; (void)s_col; (void)s_env;   // the two pixel paths share one preamble
        if (!tex_ready) { mgb_mbar_wait(&s_bar[0], 0); tex_ready = true; }
        // ---- pixels.  Each WARP owns runs of whole screen columns (c.run_px / V of them, 768 B of output): the column
        // record is warp-uniform (one broadcast read, held in registers), lanes take rows d_v = lane, lane + 32, ...,
        // write into the warp's private staging slot, and lane 0 issues ONE bulk (TMA) store per run; the slot is
        // double-buffered.  No block-wide barrier inside an env: a warp with cheap pixels (walls) simply moves on.
        const int total_px = H * V;
        const int lb_sx = trunc_i(c.lb_sx), lb_ex = s_env[4];
        const int lb_sy = trunc_i(c.lb_sy);
        int lb_ey = trunc_i(c.lb_sy + c.lb_w);
        if (lb_ey > V) lb_ey = V;
        const bool has_bar = c.task_type == MGB_MAZE_SURVIVAL;
        uint8_t *gobs = reinterpret_cast<uint8_t *>(a.obs) + (size_t)e * total_px * px_bytes;
        const int lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
        const double inv_cell = th->inv_cell, inv_t2c = th->inv_t2c;
        const bool cell_p2 = th->cell_pow2 != 0, t2c_p2 = th->t2c_pow2 != 0, text_p2 = c.text_pow2 != 0;
        const double *efft = (c.n_cls > 0 && th->cls >= 0) ? a.efftab + (size_t)th->cls * total_px : nullptr;
        const double *fogt = efft ? a.fogtab + (size_t)th->cls * total_px : nullptr;
        const double fog_from = 0.4999 * c.max_vision;
        const double dts = (double)ts;
        // Exact integer texel / cell indices.  With cell_size = 2^a, text_size = 2^b (a >= b) and ts = 2^m, every scaling in
        //   texel = trunc(ts * frac((cell/text) * frac(x / cell)))   (floor, ray_caster_utils.py:106-112)
        //   texel = trunc(ts * frac(x / text))                        (ceiling, :140-144),   cell = trunc(x / cell)
        // is by a power of two and every frac() is exact, so for x >= 0 all three equal integer functions of
        // t = trunc(x * ts / text): texel = t mod ts, cell = t >> log2(ts * cell / text).  One multiply and one conversion per
        // axis instead of three multiplies, two floors, two subtractions and two conversions; pixels with a negative
        // coordinate (the reference truncates those toward zero) and tasks with other sizes take the original expressions.
        const bool fast_ix = cell_p2 && t2c_p2 && text_p2 && (ts & (ts - 1)) == 0 && cell_size >= c.text_size;
        const double tex_scale = dts * c.inv_text;
        int cell_shift = 0;
        { int ex = 0; frexp(tex_scale * cell_size, &ex); cell_shift = ex - 1; }
        const int ts_mask = ts - 1;
        const double ix_lim = 1073741824.0 / tex_scale;                        // t < 2^30

        // floor / ceiling pixel (ray_caster_utils.py:94-153) at effective distance `eff`.  Returns true when a transparent
        // cell tinted it (`mark`).  Floor and ceiling rows share ONE instruction stream on the integer-index path (selects
        // for texture id, fog weight and tint threshold): a warp pass that holds both kinds does not run two code paths.
        auto fc_px = [&](const RowRec &rr, double eff, double fog, const ColRec &cr, int rgb[3]) -> bool {
            bool mark = false;
            if (rr.kind == 0) return false;
            const double hit_x = eff * cr.cos_abs + pos_x;
            const double hit_y = eff * cr.sin_abs + pos_y;
            const bool fastpx = fast_ix && hit_x >= 0.0 && hit_y >= 0.0 && hit_x < ix_lim && hit_y < ix_lim;
            if (fastpx) {
                const int tx = trunc_i(hit_x * tex_scale), ty = trunc_i(hit_y * tex_scale);
                const int i = tx >> cell_shift, j = ty >> cell_shift;
                const int tu = tx & ts_mask, tv_ = ty & ts_mask;
                const bool inside = (unsigned)i < (unsigned)n && (unsigned)j < (unsigned)n;
                const bool is_floor = rr.kind == 1;
                const int cell = inside ? i * n + j : 0;
                const int text_id = is_floor ? (int)texts[cell] : c.n_tex;
                // floor: 1 - alpha * light (:118), ceiling: 1 - alpha (:149); alpha * 1.0 == alpha exactly
                const double oma = 1.0 - fog * (is_floor ? rr.light : 1.0);
                shade(rgb, rr.light, oma, s_tex[(text_id * ts + tu) * ts + tv_]);
                if (is_floor && !inside) { rgb[0] = 0; rgb[1] = 0; rgb[2] = 0; }   // the floor is drawn inside the maze only (:104)
                const double tv = inside ? s_transp[cell] : 0.0;
                if (tv > (is_floor ? 0.01 : 0.0)) { blend(rgb, tv * 0.50 + 0.10); mark = true; }      // :119-123, :150-153
                return mark;
            }
            const double ci = cell_p2 ? hit_x * inv_cell : hit_x / cell_size;
            const double cj = cell_p2 ? hit_y * inv_cell : hit_y / cell_size;
            const int i = trunc_i(ci), j = trunc_i(cj);
            const bool inside = (unsigned)i < (unsigned)n && (unsigned)j < (unsigned)n;
            if (rr.kind == 1) {                                               // floor, :103-126
                if (inside) {
                    const int text_id = texts[i * n + j];
                    double d_i = ci - floor(ci), d_j = cj - floor(cj);
                    d_i = t2c_p2 ? d_i * inv_t2c : d_i / text_to_cell;
                    d_j = t2c_p2 ? d_j * inv_t2c : d_j / text_to_cell;
                    d_i -= floor(d_i); d_j -= floor(d_j);
                    d_i *= dts; d_j *= dts;
                    shade(rgb, rr.light, 1.0 - fog * rr.light, s_tex[(text_id * ts + trunc_i(d_i)) * ts + trunc_i(d_j)]);
                    const double tv = s_transp[i * n + j];
                    if (tv > 0.01) { blend(rgb, tv * 0.50 + 0.10); mark = true; }
                }
            } else {                                                          // ceiling, :137-153
                const double fi = text_p2 ? hit_x * c.inv_text : hit_x / c.text_size;
                const double fj = text_p2 ? hit_y * c.inv_text : hit_y / c.text_size;
                double d_i = fi - floor(fi), d_j = fj - floor(fj);
                d_i *= dts; d_j *= dts;
                shade(rgb, rr.light, 1.0 - fog, s_tex[(c.n_tex * ts + trunc_i(d_i)) * ts + trunc_i(d_j)]);
                if (inside) {
                    const double tv = s_transp[i * n + j];
                    if (tv > 0) { blend(rgb, tv * 0.50 + 0.10); mark = true; }
                }
            }
            return mark;
        };
        // effective distance of pixel (d_h, d_v): tabulated per height class, else the division itself (:97, :131)
        auto eff_of = [&](const double *effc, int d_v, const ColRec &cr) -> double {
            return effc ? __ldg(effc + d_v) : s_row[d_v].distance / cr.cos_hp;
        };
        // alpha = clip(2 eff / max_vision - 1, 0, 1): tabulated with eff, else evaluated here.  It is exactly 0 while
        // 2 eff / max_vision < 1; the margin keeps that shortcut independent of the division's rounding
        auto fog_of = [&](const double *fogc, int d_v, double eff) -> double {
            if (fogc) return __ldg(fogc + d_v);
            return eff > fog_from ? fmin(1.0, fmax(2.0 * eff / c.max_vision - 1.0, 0.0)) : 0.0;
        };
        // wall pixel (:184-189)
        auto wall_px = [&](int d_v, const ColRec &cr, int rgb[3]) {
            const double local_v = s_rowc[d_v] * cr.ratio + vision_height;
            double d_j = text_p2 ? local_v * c.inv_text : local_v / c.text_size;
            d_j -= floor(d_j);
            shade(rgb, cr.light, cr.oma, s_tex[(cr.text_id * ts + cr.ti) * ts + trunc_i(dts * d_j)]);
        };
        const bool out_u8 = c.obs_dtype == MGB_OBS_U8;
        auto store_px = [&](uint8_t *buf, int p, const int rgb[3]) {
            if (out_u8) {
                buf[p * 3 + 0] = (uint8_t)(rgb[0] > 255 ? 255 : rgb[0]);
                buf[p * 3 + 1] = (uint8_t)(rgb[1] > 255 ? 255 : rgb[1]);
                buf[p * 3 + 2] = (uint8_t)(rgb[2] > 255 ? 255 : rgb[2]);
            } else {
                int32_t *o = reinterpret_cast<int32_t *>(buf) + p * 3;
                o[0] = MGB_OBS_WORD(c, rgb[0]); o[1] = MGB_OBS_WORD(c, rgb[1]); o[2] = MGB_OBS_WORD(c, rgb[2]);
            }
        };

        const int cols_per_run = c.run_px / V > 0 ? c.run_px / V : 1;
        const int n_runs = (H + cols_per_run - 1) / cols_per_run;
        auto next_run = [&](int prev) {
            if (!dyn) return prev < 0 ? warp : prev + n_warps;
            int r = 0;
            if (lane == 0) r = atomicAdd(&s_runctr[b], 1);
            return __shfl_sync(0xffffffffu, r, 0);
        };
        for (int run = next_run(-1); run < n_runs; run = next_run(run)) {
            uint8_t *buf = s_out + ((size_t)warp * n_slots + (n_slots == 2 ? run_parity : 0)) * run_bytes;
            const int h0 = run * cols_per_run;
            const int ncol = H - h0 < cols_per_run ? H - h0 : cols_per_run;
            // the bulk store that last used this slot (two runs ago) must have finished reading it
            if (lane == 0) { if (n_slots == 2) mgb_bulk_wait_read<1>(); else mgb_bulk_wait_read<0>(); }
            __syncwarp();
            for (int cc = 0; cc < ncol; ++cc) {
                const int d_h = h0 + cc;
                const ColRec cr = s_col[d_h];                              // warp-uniform
                const bool bar_col = has_bar && d_h >= lb_sx && d_h < lb_ex;
                const double *effc = efft ? efft + (size_t)d_h * V : nullptr;
                const double *fogc = efft ? fogt + (size_t)d_h * V : nullptr;
                if (cr.n_hits == 0) {
                    // Column without transparent crossings: every pixel is EITHER wall OR floor/ceiling.  The two kinds are
                    // walked separately -- the wall span [ws, we), then the remaining rows as one compacted index range --
                    // so that a warp pass runs one code path with (almost) all lanes instead of both paths half empty.
                    int ws = 0, we = 0;
                    if (cr.wall) {
                        ws = cr.v_s < 0 ? 0 : (cr.v_s > V ? V : cr.v_s);
                        we = cr.v_e < ws ? ws : (cr.v_e > V ? V : cr.v_e);
                    }
                    const int span = we - ws, rest = V - span;
                    // the floor/ceiling rows' distances come from L2: ask for the first two passes' worth now, use them after
                    // the wall passes
                    double eff_a = 0.0, eff_b = 0.0, fog_a = 0.0, fog_b = 0.0;
                    {
                        const int ka = lane, kb = lane + 32;
                        if (ka < rest) {
                            const int d_v = ka < ws ? ka : ka + span;
                            eff_a = eff_of(effc, d_v, cr); fog_a = fog_of(fogc, d_v, eff_a);
                        }
                        if (kb < rest) {
                            const int d_v = kb < ws ? kb : kb + span;
                            eff_b = eff_of(effc, d_v, cr); fog_b = fog_of(fogc, d_v, eff_b);
                        }
                    }
                    for (int d_v = ws + lane; d_v < we; d_v += 32) {
                        int rgb[3];
                        wall_px(d_v, cr, rgb);
                        if (bar_col && d_v >= lb_sy && d_v < lb_ey) { rgb[0] = 255; rgb[1] = 0; rgb[2] = 0; }
                        store_px(buf, cc * V + d_v, rgb);
                    }
                    for (int k = lane; k < rest; k += 32) {
                        const int d_v = k < ws ? k : k + span;
                        int rgb[3] = {0, 0, 0};
                        const double eff = k < 32 ? eff_a : (k < 64 ? eff_b : eff_of(effc, d_v, cr));
                        const double fog = k < 32 ? fog_a : (k < 64 ? fog_b : fog_of(fogc, d_v, eff));
                        fc_px(s_row[d_v], eff, fog, cr, rgb);
                        if (bar_col && d_v >= lb_sy && d_v < lb_ey) { rgb[0] = 255; rgb[1] = 0; rgb[2] = 0; }
                        store_px(buf, cc * V + d_v, rgb);
                    }
                    continue;
                }
                const HitRec *hits = s_hit + (size_t)d_h * c.max_hits;
                for (int d_v = lane; d_v < V; d_v += 32) {
                    int rgb[3] = {0, 0, 0};
                    bool mark = false;
                    const bool in_wall = cr.wall && d_v >= cr.v_s && d_v < cr.v_e;
                    const double eff = eff_of(effc, d_v, cr);
                    mark = fc_px(s_row[d_v], eff, fog_of(fogc, d_v, eff), cr, rgb);   // also under a wall: `mark` decides about the overlays below
                    if (in_wall) wall_px(d_v, cr, rgb);
                    if (!mark) {                                              // transparent overlays, :191-205
                        for (int k = 0; k < cr.n_hits; ++k)
                            if (d_v >= hits[k].v_s && d_v < hits[k].v_e) blend(rgb, hits[k].tf);
                    }
                    if (bar_col && d_v >= lb_sy && d_v < lb_ey) {             // life bar, maze_discrete_3d.py:118-126
                        rgb[0] = 255; rgb[1] = 0; rgb[2] = 0;
                    }
                    store_px(buf, cc * V + d_v, rgb);
                }
            }
            uint8_t *dst = gobs + (size_t)h0 * V * px_bytes;
            const uint32_t bytes = (uint32_t)(ncol * V) * (uint32_t)px_bytes;
            if ((bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
                mgb_fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mgb_bulk_store(dst, buf, bytes);
                    mgb_bulk_commit();
                }
            } else {
                __syncwarp();
                for (uint32_t i = lane; i < bytes; i += 32) dst[i] = buf[i];
            }
            run_parity ^= 1;
        }
    };
    // ---- FILL: static layers of pose e (record set 0)
    auto fill_pixels = [&](int64_t e, int bb) {
        const int b = 0;
        uint8_t *s_blob = (bb ? s_blob2[1] : s_blob2[0]);
        double *s_transp = (b ? s_transp2[1] : s_transp2[0]);
        ColRec *s_col = (b ? s_col2[1] : s_col2[0]);
        RowRec *s_row = (b ? s_row2[1] : s_row2[0]);
        HitRec *s_hit = (b ? s_hit2[1] : s_hit2[0]);
        int *s_env = (b ? s_env2[1] : s_env2[0]);
        double *s_pose = (b ? s_pose2[1] : s_pose2[0]);
        const TaskHdr *th = blob_hdr(s_blob);
        const int gx = s_env[0], gy = s_env[1], ori = s_env[2], steps = s_env[3];
        const double cell_size = th->cell_size, vision_height = th->agent_height, ceil_height = th->wall_height;
        const bool cont_pose = !FILL && c.kind == MGB_MAZE_CONTINUOUS_3D;
        const double pos_x = cont_pose ? s_pose[0] : gx * cell_size + 0.5 * cell_size;   // get_cell_center, maze_base.py:194-197
        const double pos_y = cont_pose ? s_pose[1] : gy * cell_size + 0.5 * cell_size;
        const double text_to_cell = c.text_size / cell_size;

        const int8_t *walls = reinterpret_cast<const int8_t *>(s_blob + c.off_walls);
        const int8_t *texts = reinterpret_cast<const int8_t *>(s_blob + c.off_texts);
        const float *ct = a.coltab + (size_t)(cont_pose ? 0 : ori) * 3 * H;
        (void)walls; (void)steps; (void)ct; (void)ceil_height; (void)text_to_cell; (void)s_transp; (void)s_hit; (void)s_col; (void)s_env;   // the two pixel paths share one preamble
            // ---- static layers of this pose: colour before any transparency (wall colour wins inside the wall span),
            // the food slot of the floor / ceiling cell under every pixel, one word + one byte per pixel, coalesced
            const int total_px = H * V;
            const double inv_cell = th->inv_cell, inv_t2c = th->inv_t2c;
            const bool cell_p2 = th->cell_pow2 != 0, t2c_p2 = th->t2c_pow2 != 0, text_p2 = c.text_pow2 != 0;
            const double *efft = (c.n_cls > 0 && th->cls >= 0) ? a.efftab + (size_t)th->cls * total_px : nullptr;
            const double fog_from = 0.4999 * c.max_vision;
            const double dts = (double)ts;
            const int8_t *fidx = reinterpret_cast<const int8_t *>(s_blob + c.off_fidx);
            const int goal_cell = th->goal[0] * n + th->goal[1];
            uint32_t *gpx = a.c_px + (size_t)e * total_px;
            uint8_t *gfid = a.c_fid + (size_t)e * total_px;
            for (int q = tid; q < total_px; q += blockDim.x) {
                const int d_h = q / V, d_v = q - d_h * V;
                const ColRec &cr = s_col[d_h];
                const RowRec &rr = s_row[d_v];
                int rgb[3] = {0, 0, 0};
                int fid = 0xFF;
                const bool in_wall = cr.wall && d_v >= cr.v_s && d_v < cr.v_e;
                if (rr.kind != 0 && (!in_wall || cr.n_hits > 0)) {
                    const double eff = efft ? __ldg(efft + q) : rr.distance / cr.cos_hp;
                    double fog = 0.0;
                    if (eff > fog_from) fog = fmin(1.0, fmax(2.0 * eff / c.max_vision - 1.0, 0.0));
                    const double hit_x = eff * cr.cos_abs + pos_x;
                    const double hit_y = eff * cr.sin_abs + pos_y;
                    const double ci = cell_p2 ? hit_x * inv_cell : hit_x / cell_size;
                    const double cj = cell_p2 ? hit_y * inv_cell : hit_y / cell_size;
                    const int i = trunc_i(ci), j = trunc_i(cj);
                    const bool inside = (unsigned)i < (unsigned)n && (unsigned)j < (unsigned)n;
                    if (inside) {
                        if (c.task_type == MGB_MAZE_SURVIVAL) { const int f = fidx[i * n + j]; if (f >= 0) fid = f; }
                        else if (i * n + j == goal_cell) fid = 0;
                    }
                    if (rr.kind == 1) {
                        if (inside) {
                            double d_i = ci - floor(ci), d_j = cj - floor(cj);
                            const int text_id = texts[i * n + j];
                            d_i = t2c_p2 ? d_i * inv_t2c : d_i / text_to_cell;
                            d_j = t2c_p2 ? d_j * inv_t2c : d_j / text_to_cell;
                            d_i -= floor(d_i); d_j -= floor(d_j);
                            d_i *= dts; d_j *= dts;
                            shade(rgb, rr.light, 1.0 - fog * rr.light,
                                  s_tex[(text_id * ts + trunc_i(d_i)) * ts + trunc_i(d_j)]);
                        } else fid = 0xFF;
                    } else {
                        const double fi = text_p2 ? hit_x * c.inv_text : hit_x / c.text_size;
                        const double fj = text_p2 ? hit_y * c.inv_text : hit_y / c.text_size;
                        double d_i = fi - floor(fi), d_j = fj - floor(fj);
                        d_i *= dts; d_j *= dts;
                        shade(rgb, rr.light, 1.0 - fog, s_tex[(c.n_tex * ts + trunc_i(d_i)) * ts + trunc_i(d_j)]);
                    }
                }
                if (in_wall) {
                    const double local_v = (c.half_v - (d_v + 0.5) * c.pixel_size) * cr.ratio + vision_height;
                    double d_j = text_p2 ? local_v * c.inv_text : local_v / c.text_size;
                    d_j -= floor(d_j);
                    shade(rgb, cr.light, cr.oma, s_tex[(cr.text_id * ts + cr.ti) * ts + trunc_i(dts * d_j)]);
                }
                gpx[q] = (uint32_t)rgb[0] | ((uint32_t)rgb[1] << 10) | ((uint32_t)rgb[2] << 20) | (in_wall ? (1u << 30) : 0u);
                gfid[q] = (uint8_t)fid;
                uint8_t *g8 = a.c_rgb8 + ((size_t)e * total_px + q) * 3;
                g8[0] = (uint8_t)(rgb[0] > 255 ? 255 : rgb[0]);
                g8[1] = (uint8_t)(rgb[1] > 255 ? 255 : rgb[1]);
                g8[2] = (uint8_t)(rgb[2] > 255 ? 255 : rgb[2]);
            }
    };
    // ONE loop for both orders (one call site per phase keeps the kernel's code size down):
    //   sequential (FILL, or no room for two record sets): geometry(e) by the whole CTA, barrier, pixels(e), barrier;
    //   pipelined: the first kGeoThreads threads prepare env e + 1 (record set b ^ 1) and then join the other warps, which
    //   have been painting env e (record set b) since the last barrier; the first trip only prepares.
    const bool pipe = !FILL && c.pipe;
    const int64_t stride = gridDim.x;
    int b = 0, bb = 0;
    if (!pipe && tid == 0 && (int64_t)blockIdx.x < a.n) load_blob(blockIdx.x, 0);
    for (int64_t e_pix = pipe ? (int64_t)blockIdx.x - stride : (int64_t)blockIdx.x; e_pix < a.n; e_pix += stride) {
        const int64_t e_geo = pipe ? e_pix + stride : e_pix;
        if (e_geo < a.n && (!pipe || tid < kGeoThreads)) {
            // sequential: the other tile buffer was last read before the barrier that closed the previous env: prefetch into it
            geometry(e_geo, pipe ? b ^ 1 : 0, pipe ? b ^ 1 : bb, tid, pipe ? kGeoThreads : (int)blockDim.x, pipe, pipe,
                     (!pipe && e_geo + stride < a.n) ? e_geo + stride : -1);
        }
        if (!pipe) {
            if (!tex_ready) { mgb_mbar_wait(&s_bar[0], 0); tex_ready = true; }
            __syncthreads();
        }
        if (e_pix >= 0) {
            if (FILL) fill_pixels(e_pix, bb);
            else pixels(e_pix, pipe ? b : 0, pipe ? b : bb, pipe);
        }
        __syncthreads();   // the record set / tile just painted from is rewritten next
        b ^= 1; bb ^= 1;   // pipelined: the set prepared in this trip is painted in the next one
    }
    if (!tex_ready && tid == 0) mgb_mbar_wait(&s_bar[0], 0);   // never leave a TMA load in flight
    if ((tid & 31) == 0) mgb_bulk_wait_read<0>();   // smem must outlive the copies; the kernel boundary flushes the writes
}

// What the compose step needs from an env after its step logic: pose slot, life bar, food presence, and the 8-bit
// signature of the foods that can show in this pose but are currently missing.
__device__ __forceinline__ EnvDyn make_dyn(const MazeConst &c, const MazeArgs &a, const uint8_t *blob, int task, const Env &s,
                                           const int32_t *eaten)
{
    const TaskHdr *th = blob_hdr(blob);
    EnvDyn d;
    const size_t pose = (size_t)task * c.n * c.n * 4 + (s.gx * c.n + s.gy) * 4 + s.ori;
    uint64_t fm[2];
    int vb = -1;
    bool have_vb = false;
    if (a.pose_rec) {                  // one record: nothing below waits for a second round trip
        const int4 *r = reinterpret_cast<const int4 *>(reinterpret_cast<const PoseRec *>(a.pose_rec) + pose);
        const int4 r0 = __ldg(r), r1 = __ldg(r + 1);
        d.slot = r0.x; vb = r0.y; have_vb = true;
        fm[0] = ((uint64_t)(uint32_t)r0.w << 32) | (uint32_t)r0.z;
        fm[1] = ((uint64_t)(uint32_t)r1.y << 32) | (uint32_t)r1.x;
    } else {
        d.slot = a.pose_index[pose];
    }
    d.task = task; d.pad = 0;
    d.present[0] = d.present[1] = 0;
    if (c.task_type == MGB_MAZE_SURVIVAL) {
        const int32_t *fint = reinterpret_cast<const int32_t *>(blob + c.off_fint);
        // batches of 8 foods: the 16 loads of a batch are issued together (one memory latency per batch, not per food)
        const int n_food = th->n_food;
        for (int f0 = 0; f0 < n_food; f0 += 8) {
            int ea[8], fi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int f = f0 + k < n_food ? f0 + k : n_food - 1;
                ea[k] = eaten[f * a.n_pad];
                fi[k] = fint[f];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int f = f0 + k;
                if (f < n_food && ((ea[k] == kNever) || (s.steps >= ea[k] + fi[k]))) d.present[f >> 6] |= 1ull << (f & 63);
            }
        }
        int ex = trunc_i(c.lb_sx + s.life / th->max_life * c.lb_l);      // maze_discrete_3d.py:118-126
        if (ex < 0) { ex += c.res_h; if (ex < 0) ex = 0; }
        if (ex > c.res_h) ex = c.res_h;
        d.bar_end = ex;
    } else {
        d.present[0] = 1ull;             // the goal cell is pseudo food slot 0, always present (maze_base.py:59-60)
        d.bar_end = 0;
    }
    // 8-bit signature of the foods that can show in this pose but are currently missing; 0 -> the whole frame is the
    // baked "all present" frame + life bar
    if (!a.pose_rec) { fm[0] = a.c_fmask[(size_t)d.slot * 2]; fm[1] = a.c_fmask[(size_t)d.slot * 2 + 1]; }
    uint64_t miss = ((~d.present[0]) & fm[0]) | ((~d.present[1]) & fm[1]);      // fold 128 slots to (f & 7)
    d.vframe = -1; d.pad2 = 0;
    if (miss && (have_vb || a.c_vbase)) {
        if (!have_vb) vb = a.c_vbase[d.slot];
        if (vb >= 0) {
            // variant index = presence bits of the pose's foods, compacted in ascending slot order (all visible -> the
            // c_rgb8 frame, handled by miss == 0 above)
            int v = 0, bit = 0;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                uint64_t m = fm[w];
                while (m) {
                    const int f = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    v |= (int)((d.present[w] >> f) & 1ull) << bit;
                    ++bit;
                }
            }
            d.vframe = vb + v;
            miss = 0;                    // the variant frame is final
        }
    }
    miss |= miss >> 32; miss |= miss >> 16; miss |= miss >> 8;
    d.pad = (int32_t)(miss & 0xFFu);
    return d;
}

// ---------------------------------------------------------------------------------------------------------------
// Pose cache path: step logic (one thread per env) + compose (static pose layers x current food state -> observation)
// ---------------------------------------------------------------------------------------------------------------
// DYN = false: step logic only, ahead of the direct renderer (which then runs with do_step = 0): the per-env logic is a
// chain of dependent L2 round trips that one thread of a 512-thread CTA would otherwise walk once per env, serially.
template <bool DYN>
__global__ void maze3d_logic_kernel(const __grid_constant__ MazeConst c, const __grid_constant__ MazeArgs a)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    const int task = a.env2task[e];
    const uint8_t *blob = a.blobs + (int64_t)task * c.blob_bytes;
    const TaskHdr *th = blob_hdr(blob);
    const int4 ag = a.agent[e];
    Env s = {ag.x, ag.y, ag.z, ag.w, a.life[e]};
    int32_t *eaten = a.eaten + e;
    if (a.do_step) {
        double reward;
        int done;
        maze_logic(c, blob, eaten, a.n_pad, s, a.act[e], reward, done);
        a.rew[e] = reward;
        a.done[e] = (uint8_t)done;
        if (done && a.auto_reset) env_reset(c, blob, eaten, a.n_pad, s);
        a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
        a.life[e] = s.life;
    }
    if (DYN) {
        const EnvDyn d = make_dyn(c, a, blob, task, s, eaten);
        reinterpret_cast<EnvDyn *>(a.dyn)[e] = d;
    }
}

// per pose slot, after the FILL render: (1) c_fmask = the food slots that can change this pose's image at all (under a
// pixel or in a crossing record); (2) c_gsig = one byte per 4-pixel group, OR of 1 << (f & 7) over the food slots f that
// can tint a pixel of the group (floor/ceiling cell under it, or a crossing span over it).  0 = never tinted.  An env
// whose missing foods have the 8-bit signature S needs the float64 path only for groups with (gsig & S) != 0.
__global__ void __launch_bounds__(256) maze3d_sig_kernel(const __grid_constant__ MazeConst c,
                                                         const __grid_constant__ MazeArgs a)
{
    const int64_t slot = blockIdx.x;
    const int V = c.res_v, total_px = c.res_h * V;
    const uint8_t *gfid = a.c_fid + (size_t)slot * total_px;
    const uint8_t *colhits = a.c_colhits + (size_t)slot * c.res_h;
    const HitRec *ghits = reinterpret_cast<const HitRec *>(a.c_hits) + (size_t)slot * c.res_h * c.max_hits;
    uint8_t *gsig = a.c_gsig + (size_t)slot * (total_px / 4);
    uint64_t m0 = 0, m1 = 0;
    auto note = [&](int f, uint32_t &sig) {
        if (f < 0 || f >= 128) return;
        sig |= 1u << (f & 7);
        if (f < 64) m0 |= 1ull << f; else m1 |= 1ull << (f - 64);
    };
    for (int g = threadIdx.x; g < total_px / 4; g += blockDim.x) {
        uint32_t sig = 0;
        for (int k = 0; k < 4; ++k) {
            const int q = 4 * g + k, d_h = q / V, d_v = q - d_h * V;
            const int f = gfid[q];
            if (f != 0xFF) note(f, sig);
            const HitRec *hh = ghits + (size_t)d_h * c.max_hits;
            for (int j = 0; j < colhits[d_h]; ++j)
                if (d_v >= hh[j].v_s && d_v < hh[j].v_e) note(hh[j].fid, sig);
        }
        gsig[g] = (uint8_t)sig;
    }
    for (int o = 16; o > 0; o >>= 1) { m0 |= __shfl_xor_sync(0xffffffffu, m0, o); m1 |= __shfl_xor_sync(0xffffffffu, m1, o); }
    if ((threadIdx.x & 31) == 0) {
        if (m0) atomicOr(reinterpret_cast<unsigned long long *>(a.c_fmask + slot * 2), (unsigned long long)m0);
        if (m1) atomicOr(reinterpret_cast<unsigned long long *>(a.c_fmask + slot * 2 + 1), (unsigned long long)m1);
    }
}

// variant frames start as copies of the pose's STATIC frame (what the FILL render left in c_rgb8, before any tint is baked)
__global__ void __launch_bounds__(256) maze3d_varinit_kernel(const __grid_constant__ MazeConst c, const __grid_constant__ MazeArgs a)
{
    const size_t frame16 = (size_t)c.res_h * c.res_v * 3 / 16;
    const BakeDesc bd = reinterpret_cast<const BakeDesc *>(a.bake_desc)[blockIdx.x];
    const uint4 *src = reinterpret_cast<const uint4 *>(a.c_rgb8) + (size_t)bd.slot * frame16;
    uint4 *dst = reinterpret_cast<uint4 *>(a.c_var8) + (size_t)blockIdx.x * frame16;
    for (size_t i = threadIdx.x; i < frame16; i += blockDim.x) dst[i] = src[i];
}

constexpr int kComposeThreads = 256;

// Work item = (env, image slice).  Pixels no missing food can tint are copied from the pose's baked all-present frame;
// the others take the cached static layers (packed colour + in-wall flag, food slot under the pixel, crossing records
// of the column) through the reference's float64 blend with the env's 128-bit food presence mask.
__global__ void __launch_bounds__(kComposeThreads, 5) maze3d_compose_kernel(const __grid_constant__ MazeConst c,
                                                                         const __grid_constant__ MazeArgs a)
{
    // persistent CTAs: work item = (env, one of kParts slices of its image); grid = resident CTA count, so there is
    // no partial last wave (1024 envs as 1024 CTAs ran 1.15 waves = almost twice the time of one)
    const int kParts = a.do_parts;      // 1..4 image slices per env, chosen by the host so that items >> resident CTAs
    const int H = c.res_h, V = c.res_v, total_px = H * V;
    const bool survival = c.task_type == MGB_MAZE_SURVIVAL;
    const int lb_sx = trunc_i(c.lb_sx), lb_sy = trunc_i(c.lb_sy);
    int lb_ey = trunc_i(c.lb_sy + c.lb_w);
    if (lb_ey > V) lb_ey = V;
    const int px_bytes = c.obs_dtype == MGB_OBS_U8 ? 3 : 12;
    const int part_px = ((total_px / kParts) + 127) / 128 * 128;          // slice boundaries stay 128-pixel aligned
  const int64_t n_items = a.n * kParts;
  // bake mode (pose-cache build): item = pose slot, every food present, no life bar, tintable groups only
  auto fetch = [&](int64_t idx) -> EnvDyn {
      if (!a.bake) return reinterpret_cast<const EnvDyn *>(a.dyn)[idx];
      EnvDyn b;
      b.slot = (int32_t)idx; b.bar_end = 0; b.present[0] = b.present[1] = ~0ull; b.pad = 0xFF; b.vframe = -1; b.pad2 = 0;
      if (a.bake_desc) {               // variant frames: item = frame, its pose slot and presence mask come from the descriptor
          const BakeDesc bd = reinterpret_cast<const BakeDesc *>(a.bake_desc)[idx];
          b.slot = bd.slot; b.present[0] = bd.present[0]; b.present[1] = bd.present[1];
      }
      b.task = a.poses[b.slot].x;
      return b;
  };
  EnvDyn d_next;
  if ((int64_t)blockIdx.x < n_items) d_next = fetch(blockIdx.x / kParts);
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t e = item / kParts;
    const int part = (int)(item - e * kParts);
    const int q_begin = part * part_px, q_end = (part + 1) * part_px < total_px ? (part + 1) * part_px : total_px;
    // software pipeline: the next item's EnvDyn (-> pose slot -> every address below) is requested now, so the
    // dependent-load bubble at the start of an item overlaps this item's pixels
    const EnvDyn d = d_next;
    if (item + gridDim.x < n_items) d_next = fetch((item + gridDim.x) / kParts);
    if (q_begin >= total_px) continue;
#include "maze_compose_body.inc"
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused single-step kernel of the pose-cache path (uint8 frames): step logic + observation in ONE launch, the frame moved by
// the TMA engine.  One CTA per env: thread 0 runs the step logic (action, evaluation rule, auto-reset, pose-cache lookup) and
// immediately issues bulk copies (cp.async.bulk + mbarrier, 12 KB chunks) of the pose's baked all-present frame into shared
// memory; as each chunk lands the CTA patches, in shared memory, the few 4-pixel groups a currently missing food can tint
// (float64 blend of the cached static layers, exactly the compose kernel's slow path) and the life bar, and one bulk store
// sends the chunk to `obs`.  The 49 KB of a frame are in flight without passing through registers (the compose kernel's 256
// threads of dependent 16-byte loads were 48 % long-scoreboard stalls), the separate logic launch and its EnvDyn round trip
// through global memory are gone, and four CTAs per SM overlap one env's logic latency with the others' copies.
// Reference: maze_discrete_3d.py:51-81,113-127, maze_base.py:65-95, ray_caster_utils.py:66-209 (via the cached layers).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStepThreads = 128;
constexpr int kStepChunkPx = 4096;          // pixels per bulk copy: 12 KB of uint8 RGB

// one tinted 4-pixel group (4 consecutive rows of ONE screen column): cached static colour -> floor/ceiling tint ->
// crossings of the column -> life bar (ray_caster_utils.py:118-205), packed to 12 uint8 (values above 255 clamp: MGB_OBS_U8).
// Each crossing record of the column is loaded once for the four pixels (they share the column), in ascending order, so
// every pixel sees the blends in the reference's order.
__device__ __forceinline__ void compose_group_u8(const MazeConst &c, const EnvDyn &d, const uint32_t *gpx, const uint8_t *gfid,
                                                 const uint8_t *colhits, const HitRec *ghits, const double *fval,
                                                 bool survival, int q, int d_h, int d_v0, int lb_sx, int lb_sy, int lb_ey,
                                                 uint32_t pk[3])
{
    const int V = c.res_v;
    const uint4 w4 = __ldg(reinterpret_cast<const uint4 *>(gpx + q));
    const uint32_t f4 = __ldg(reinterpret_cast<const uint32_t *>(gfid + q));
    const int n_hits = colhits[d_h];
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    int rgb[4][3];
    bool mark[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int d_v = d_v0 + k, f = (int)((f4 >> (8 * k)) & 0xFFu);
        rgb[k][0] = (int)(w[k] & 1023u); rgb[k][1] = (int)((w[k] >> 10) & 1023u); rgb[k][2] = (int)((w[k] >> 20) & 1023u);
        const bool in_wall = (w[k] >> 30) & 1u;
        mark[k] = false;
        if (f != 0xFF && ((d.present[f >> 6] >> (f & 63)) & 1ull)) {
            const double tv = survival ? __ldg(fval + f) : 1.0;
            if (d_v > V / 2 ? tv > 0.01 : tv > 0) {          // floor tests > 0.01 (:119), ceiling > 0 (:150)
                if (!in_wall) blend(rgb[k], tv * 0.50 + 0.10);
                mark[k] = true;
            }
        }
    }
    if (n_hits > 0 && !(mark[0] && mark[1] && mark[2] && mark[3])) {
        const int4 *hits = reinterpret_cast<const int4 *>(ghits + (size_t)d_h * c.max_hits);
        for (int j0 = 0; j0 < n_hits; j0 += 4) {                      // four records per round trip
            int4 raw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) raw[u] = __ldg(hits + (j0 + u < n_hits ? j0 + u : n_hits - 1));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u >= n_hits) break;
                const double tf = __hiloint2double(raw[u].y, raw[u].x);
                const int v_s = (int)(int16_t)(raw[u].z & 0xFFFF), v_e = (int)(int16_t)((uint32_t)raw[u].z >> 16), fid = raw[u].w;
                if (!((d.present[fid >> 6] >> (fid & 63)) & 1ull)) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!mark[k] && d_v0 + k >= v_s && d_v0 + k < v_e) blend(rgb[k], tf);
            }
        }
    }
    pk[0] = pk[1] = pk[2] = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int d_v = d_v0 + k;
        if (survival && d_h >= lb_sx && d_h < d.bar_end && d_v >= lb_sy && d_v < lb_ey) { rgb[k][0] = 255; rgb[k][1] = 0; rgb[k][2] = 0; }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int idx = 3 * k + b;
            pk[idx >> 2] |= (uint32_t)(rgb[k][b] > 255 ? 255 : rgb[k][b]) << (8 * (idx & 3));
        }
    }
}

constexpr int kStepBatch = 16;              // envs whose step logic one CTA runs side by side before moving their frames
constexpr int kStepSlots = 8;               // 12 KB chunk slots of a CTA's shared-memory ring
constexpr int kStepSlack = 3;               // bulk stores that may still be reading their slot before it is handed back

__global__ void __launch_bounds__(kStepThreads, 2) maze3d_step_kernel(const __grid_constant__ MazeConst c,
                                                                      const __grid_constant__ MazeArgs a)
{
    // Two CTAs per SM; CTA j owns the envs j, j + G, j + 2G, ... (G = grid size).  Per pass of up to kStepBatch envs:
    //  (1) threads 0..B-1 run the step logic of the B envs SIDE BY SIDE (each is a chain of ~4 dependent L2 round trips:
    //      one chain per frame would cost more than the frame's copy), leaving B EnvDyn records + frame pointers in shared memory;
    //  (2) the B frames stream, 12 KB chunk by chunk, through a ring of kStepSlots shared-memory slots.  Warp 1's lane 0 only
    //      ISSUES bulk loads, up to kStepSlots chunks ahead, as the consumer hands slots back through their `empty`
    //      mbarriers; warp 0 only consumes: waits for chunk u, draws the life bar into it, sends it to `obs` with one bulk
    //      store and -- once the store of chunk u - kStepSlack has left its slot -- releases that slot.  Loads and stores stay
    //      in flight across frame boundaries.  Warps 0, 2, 3 (named barrier 1, 96 threads) tint the rare frames that still need
    //      float64 blends (poses whose image depends on more foods than have variant frames).
    // The frame move is HBM-bound: scripts/microbench/framecopy.cu moves the same 1024 x 48 KB in 16.4 us (6.1 TB/s read + write)
    // with ANY scheme -- this ring, LDG/STG.128, cudaMemcpy -- so the step costs launch gap + logic (~4 us) + that.
    extern __shared__ __align__(128) uint8_t s_ring[];           // kStepSlots x 12 KB
    __shared__ EnvDyn s_dyn[kStepBatch];
    __shared__ const uint8_t *s_src[kStepBatch];                 // finished frame each env's observation starts from
    __shared__ __align__(8) uint64_t s_bar[kStepSlots];
    __shared__ __align__(8) uint64_t s_empty[kStepSlots];
    __shared__ int s_nslow;
    __shared__ uint16_t s_slow[1024];                            // queued tinted groups of one chunk (group index in the chunk)
    const int H = c.res_h, V = c.res_v, total_px = H * V;
    const uint32_t frame_bytes = (uint32_t)total_px * 3u;
    const int n_chunks = (total_px + kStepChunkPx - 1) / kStepChunkPx;
    const bool survival = c.task_type == MGB_MAZE_SURVIVAL;
    const int lb_sx = trunc_i(c.lb_sx), lb_sy = trunc_i(c.lb_sy);
    int lb_ey = trunc_i(c.lb_sy + c.lb_w);
    if (lb_ey > V) lb_ey = V;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool warp0 = tid < 32;
    const int v_shift = (V & (V - 1)) == 0 ? 31 - __clz(V) : -1;
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        for (int k = 0; k < kStepSlots; ++k) { mgb_mbar_init(&s_bar[k], 1); mgb_mbar_init(&s_empty[k], 1); }
        mgb_fence_mbar_init();
        s_nslow = 0;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    uint32_t g0 = 0;                                             // chunks this CTA has moved in earlier passes (ring position)
    for (int64_t base = blockIdx.x; base < a.n; base += (int64_t)gridDim.x * kStepBatch) {
        const int64_t left = (a.n - base + gridDim.x - 1) / gridDim.x;
        const int B = (int)(left < kStepBatch ? left : kStepBatch);
        const int M = B * n_chunks;                               // chunks of this pass
        // ---- (1) step logic of the pass's envs, one thread each (what maze3d_logic_kernel does for the two-kernel path)
        if (tid < B) {
            const int64_t e = base + (int64_t)tid * gridDim.x;
            const int task = a.env2task[e];
            const uint8_t *blob = a.blobs + (int64_t)task * c.blob_bytes;
            const int4 ag = a.agent[e];
            Env s = {ag.x, ag.y, ag.z, ag.w, a.life[e]};
            int32_t *eaten = a.eaten + e;
            if (a.do_step) {
                double reward;
                int done;
                maze_logic(c, blob, eaten, a.n_pad, s, a.act[e], reward, done);
                a.rew[e] = reward;
                a.done[e] = (uint8_t)done;
                if (done && a.auto_reset) env_reset(c, blob, eaten, a.n_pad, s);
                a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
                a.life[e] = s.life;
            }
            const EnvDyn dd = make_dyn(c, a, blob, task, s, eaten);
            s_dyn[tid] = dd;
            s_src[tid] = frame_of(a, dd, frame_bytes);
        }
        __syncthreads();
        const bool producer = (tid >> 5) == 1;
        const int t96 = tid < 32 ? tid : tid - 32;                 // index inside the consumer group (warps 0, 2, 3)
        if (producer) {
            if (lane == 0) {
                int it = 0, k = 0;
                for (int u = 0; u < M; ++u) {
                    const uint32_t gu = g0 + (uint32_t)u;
                    const int slot = (int)(gu % kStepSlots);
                    const uint32_t use = gu / kStepSlots;           // how often this slot has been filled before
                    if (use > 0) mgb_mbar_wait(&s_empty[slot], (use - 1) & 1u);
                    const uint32_t off = (uint32_t)k * kStepChunkPx * 3u;
                    const uint32_t bytes = frame_bytes - off < kStepChunkPx * 3u ? frame_bytes - off : kStepChunkPx * 3u;
                    mgb_mbar_expect_tx(&s_bar[slot], bytes);
                    mgb_bulk_load(s_ring + (size_t)slot * (kStepChunkPx * 3), s_src[it] + off, bytes, &s_bar[slot]);
                    if (++k == n_chunks) { k = 0; ++it; }
                }
            }
        } else {
        for (int it = 0; it < B; ++it) {
            asm volatile("bar.sync 1, 96;" ::: "memory");          // frame boundary: nobody runs more than a frame ahead of warp 0
            const int64_t e = base + (int64_t)it * gridDim.x;
            const EnvDyn d = s_dyn[it];
            const uint32_t miss_sig = (uint32_t)d.pad & 0xFFu;     // != 0: some groups need the float64 path (uniform)
            uint8_t *gobs = reinterpret_cast<uint8_t *>(a.obs) + (size_t)e * frame_bytes;
            if (!warp0 && !miss_sig) continue;
            const uint8_t *blob = a.blobs + (int64_t)d.task * c.blob_bytes;
            const double *fval = reinterpret_cast<const double *>(blob + c.off_fval);
            const uint32_t *gpx = a.c_px + (size_t)d.slot * total_px;
            const uint8_t *gfid = a.c_fid + (size_t)d.slot * total_px;
            const uint8_t *colhits = a.c_colhits + (size_t)d.slot * H;
            const HitRec *ghits = reinterpret_cast<const HitRec *>(a.c_hits) + (size_t)d.slot * H * c.max_hits;
            const uint8_t *gsig = a.c_gsig + (size_t)d.slot * (total_px / 4);
            const uint32_t miss4 = miss_sig * 0x01010101u;
            for (int k = 0; k < n_chunks; ++k) {
                const int u = it * n_chunks + k;
                const uint32_t gu = g0 + (uint32_t)u;
                const int slot = (int)(gu % kStepSlots);
                const uint32_t parity = (gu / kStepSlots) & 1u;
                uint8_t *s_chunk = s_ring + (size_t)slot * (kStepChunkPx * 3);
                const int q0 = k * kStepChunkPx, q1 = q0 + kStepChunkPx < total_px ? q0 + kStepChunkPx : total_px;
                const uint32_t off = (uint32_t)q0 * 3u, bytes = (uint32_t)(q1 - q0) * 3u;
                if (miss_sig) {
                    // whole CTA: queue the chunk's tinted groups (they cluster in a few columns) ...
                    for (int g4 = (q0 >> 4) + t96; g4 < (q1 >> 4); g4 += 96) {
                        uint32_t hit = __ldg(reinterpret_cast<const uint32_t *>(gsig) + g4) & miss4;
                        while (hit) {
                            const int g = (__ffs(hit) - 1) >> 3;
                            hit &= ~(0xFFu << (8 * g));
                            s_slow[atomicAdd(&s_nslow, 1)] = (uint16_t)(((g4 << 2) + g) - (q0 >> 2));
                        }
                    }
                    asm volatile("bar.sync 1, 96;" ::: "memory");
                    const int n_slow = s_nslow;
                    mgb_mbar_wait(&s_bar[slot], parity);
                    // ... and share them out evenly: float64 blends of the cached static layers over the baked pixels
                    for (int i = t96; i < n_slow; i += 96) {
                        const int q = q0 + 4 * (int)s_slow[i];
                        const int d_h = v_shift >= 0 ? (q >> v_shift) : q / V;
                        uint32_t pk[3];
                        compose_group_u8(c, d, gpx, gfid, colhits, ghits, fval, survival, q, d_h, q - d_h * V, lb_sx, lb_sy, lb_ey, pk);
                        uint32_t *dst = reinterpret_cast<uint32_t *>(s_chunk + (size_t)(q - q0) * 3);
                        dst[0] = pk[0]; dst[1] = pk[1]; dst[2] = pk[2];
                    }
                    mgb_fence_proxy_async();
                    asm volatile("bar.sync 1, 96;" ::: "memory");  // tints done before the bar is drawn over them / the store
                    if (tid == 0) s_nslow = 0;
                }
                if (warp0) {
                    if (!miss_sig) mgb_mbar_wait(&s_bar[slot], parity);
                    // life bar columns inside this chunk (drawn last, maze_discrete_3d.py:118-126)
                    const int h0 = q0 / V, h1 = (q1 + V - 1) / V;      // columns [h0, h1) intersect the chunk
                    const int bh0 = lb_sx > h0 ? lb_sx : h0, bh1 = d.bar_end < h1 ? d.bar_end : h1;
                    if (survival && bh0 < bh1 && lb_sy < lb_ey) {
                        // one lane per column, rows walked in order: no integer division on the warp that feeds the ring
                        for (int d_h = bh0 + lane; d_h < bh1; d_h += 32) {
                            const int qc = d_h * V - q0;                // the column's first pixel, relative to the chunk
                            int v0 = lb_sy, v1 = lb_ey;                 // rows whose pixel lies inside [q0, q1)
                            if (qc + v0 < 0) v0 = -qc;
                            if (qc + v1 > q1 - q0) v1 = q1 - q0 - qc;
                            uint8_t *px = s_chunk + (size_t)(qc + v0) * 3;
                            for (int d_v = v0; d_v < v1; ++d_v, px += 3) { px[0] = 255; px[1] = 0; px[2] = 0; }
                        }
                        mgb_fence_proxy_async();
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mgb_bulk_store(gobs + off, s_chunk, bytes);
                        mgb_bulk_commit();
                        if (u >= kStepSlack) {
                            mgb_bulk_wait_read<kStepSlack>();      // the store of chunk u - 3 has left its slot: hand it back
                            mgb_mbar_arrive(&s_empty[(int)((gu - kStepSlack) % kStepSlots)]);
                        }
                    }
                }
            }
        }
        if (tid == 0) {                                            // the pass's last stores: wait for them, release their slots
            mgb_bulk_wait_read<0>();
            for (int u = (M > kStepSlack ? M - kStepSlack : 0); u < M; ++u)
                mgb_mbar_arrive(&s_empty[(int)((g0 + (uint32_t)u) % kStepSlots)]);
        }
        }
        g0 += (uint32_t)M;
        __syncthreads();                                           // s_dyn is rewritten by the next pass
    }
    if (tid == 0) mgb_bulk_wait_read<0>();       // shared memory must outlive the copies
}

// T MetaMazeDiscrete3D steps in one launch (pose-cache path): one CTA per env; thread 0 runs the step logic and leaves the
// env's EnvDyn in shared memory, then the whole CTA composes frame t straight into obs[t][env].  No logic launch, no
// EnvDyn round trip through global memory, and the logic of one env overlaps the pixels of the others on the SM.
__global__ void __launch_bounds__(kComposeThreads, 5) maze3d_rollout_kernel(const __grid_constant__ MazeConst c,
                                                                            const __grid_constant__ MazeArgs a)
{
    __shared__ EnvDyn s_dyn;
    __shared__ int s_nslow;
    extern __shared__ int s_slow[];                     // total_px / 4 entries: queued tinted groups of the current frame
    bool deferred_pass = false;
    const int H = c.res_h, V = c.res_v, total_px = H * V;
    const bool survival = c.task_type == MGB_MAZE_SURVIVAL;
    const int lb_sx = trunc_i(c.lb_sx), lb_sy = trunc_i(c.lb_sy);
    int lb_ey = trunc_i(c.lb_sy + c.lb_w);
    if (lb_ey > V) lb_ey = V;
    const int px_bytes = c.obs_dtype == MGB_OBS_U8 ? 3 : 12;
    const uint2 akey = make_uint2((uint32_t)a.act_seed, (uint32_t)(a.act_seed >> 32));
    for (int64_t env = blockIdx.x; env < a.n; env += gridDim.x) {
        const int task = a.env2task[env];
        const uint8_t *eblob = a.blobs + (int64_t)task * c.blob_bytes;
        int32_t *eaten = a.eaten + env;
        Env s = {0, 0, 0, 0, 0.0};
        if (threadIdx.x == 0) {
            const int4 ag = a.agent[env];
            s.gx = ag.x; s.gy = ag.y; s.ori = ag.z; s.steps = ag.w; s.life = a.life[env];
        }
        const int64_t genv = a.env_base + env;
        for (int t = 0; t < a.T; ++t) {
            if (threadIdx.x == 0) {
                int action;
                if (a.act) action = a.act[(int64_t)t * a.n + env];
                else {
                    const uint4 r = mgb_philox4x32_10(make_uint4((uint32_t)genv, (uint32_t)((uint64_t)genv >> 32),
                                                                 a.t_base + (uint32_t)t, MGB_STREAM_ACTION), akey);
                    action = (int)(r.x >> 30);
                    if (a.act_out) a.act_out[(int64_t)t * a.n + env] = action;
                }
                double reward;
                int done;
                maze_logic(c, eblob, eaten, a.n_pad, s, action, reward, done);
                if (done && a.auto_reset) env_reset(c, eblob, eaten, a.n_pad, s);
                if (a.rew) a.rew[(int64_t)t * a.n + env] = reward;
                if (a.done) a.done[(int64_t)t * a.n + env] = (uint8_t)done;
                s_dyn = make_dyn(c, a, eblob, task, s, eaten);
                s_nslow = 0;
            }
            __syncthreads();
            if (a.obs) {
                const EnvDyn d = s_dyn;
                const int64_t e = (int64_t)t * a.n + env;             // frame index of the included body
                const int q_begin = 0, q_end = total_px;
#define MGB_COMPOSE_DEFER_SLOW 1
#include "maze_compose_body.inc"
#undef MGB_COMPOSE_DEFER_SLOW
            }
            __syncthreads();                                          // s_dyn is rewritten by the next step
        }
        if (threadIdx.x == 0) {
            a.agent[env] = make_int4(s.gx, s.gy, s.ori, s.steps);
            a.life[env] = s.life;
        }
    }
}

// Pose-independent part of the floor/ceiling geometry: eff(d_h, d_v) = distance(d_v) / cos_hp(d_h)
// (ray_caster_utils.py:97-104,131-138) depends only on the screen, the optics and the two heights of a task, so it is
// tabulated once per (agent_height, wall_height) class with the SAME float64 division the renderer would execute.
__global__ void maze_efftab_kernel(const __grid_constant__ MazeConst c, const float *coltab, const double *cls_heights,
                                   int n_cls, double *efftab, double *fogtab)
{
    const int H = c.res_h, V = c.res_v;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_cls * H * V) return;
    const int k = (int)(idx / ((int64_t)H * V));
    const int q = (int)(idx - (int64_t)k * H * V);
    const int d_h = q / V, d_v = q - d_h * V;
    const double vision_height = cls_heights[2 * k], ceil_height = cls_heights[2 * k + 1];
    double distance = 0.0;
    if (d_v > V / 2) {
        const double v_screen = (d_v + 0.5) * c.pixel_size - c.half_v;
        distance = vision_height / v_screen * c.l_focal;
    } else if (d_v < V / 2) {
        const double v_screen = c.half_v - (d_v + 0.5) * c.pixel_size;
        distance = (ceil_height - vision_height) / v_screen * c.l_focal;
    }
    const double eff = distance / (double)coltab[d_h];     // cos_hp does not depend on the heading: use heading 0
    efftab[idx] = eff;
    // the fog weight of that distance (ray_caster_utils.py:99,133), the expression the renderer evaluates when it has no table
    fogtab[idx] = eff > 0.4999 * c.max_vision ? fmin(1.0, fmax(2.0 * eff / c.max_vision - 1.0, 0.0)) : 0.0;
}

__global__ void maze_state_kernel(MazeArgs a, int32_t *agent_out, double *life_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    const int4 ag = a.agent[e];
    agent_out[4 * e + 0] = ag.x; agent_out[4 * e + 1] = ag.y; agent_out[4 * e + 2] = ag.z; agent_out[4 * e + 3] = ag.w;
    if (life_out) life_out[e] = a.life[e];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// handle + C ABI
// ---------------------------------------------------------------------------------------------------------------
struct mgb_maze {
    int device = 0;
    int64_t n = 0, n_pad = 0, env_base = 0;
    mgb_maze_cfg cfg;
    MazeConst c;
    int4 *agent = nullptr;
    double *life = nullptr;
    int32_t *eaten = nullptr;
    int32_t *env2task = nullptr;
    uint8_t *blobs = nullptr;
    uint32_t *tex = nullptr;
    float *coltab = nullptr;
    double *efftab = nullptr, *fogtab = nullptr;
    float2 *cpos = nullptr;        // continuous maze pose
    double *cori = nullptr;
    double *coltab_d = nullptr;
    // pose cache
    int step_pdl = 0;              // MGB_MAZE_PDL=1: programmatic dependent launch between consecutive fused steps (measured 1 % slower)
    int render_pipe = 1;           // MGB_MAZE_RENDER_PIPE=0: direct renderer without the geometry / pixel software pipeline
    int fused_step = 1;            // MGB_MAZE_FUSED_STEP=0: logic kernel + compose kernel instead of maze3d_step_kernel
    size_t step_smem_set = 0, m2d_smem_set = 0;
    int cache_enabled = 1;         // MGB_MAZE_CACHE=0 disables (direct renderer only)
    double cache_budget_gb = 24.0; // MGB_MAZE_CACHE_GB
    bool cache_ready = false, cache_dirty = true;
    int64_t n_poses = 0;
    int4 *poses = nullptr;
    int32_t *pose_index = nullptr;
    uint32_t *c_px = nullptr;
    uint8_t *c_fid = nullptr, *c_colhits = nullptr, *c_rgb8 = nullptr;
    uint32_t *c_px_all = nullptr;
    uint64_t *c_fmask = nullptr;
    uint32_t *task_epoch = nullptr;           // [n_pad] how often each env's task has been resampled on the device
    bool slot_per_env = false;                // env2task is injective: every env owns its task-table slot
    uint8_t *task_flags = nullptr;            // [n_tasks] scratch of mgb_maze_update_tasks
    int task_flags_n = 0;
    bool cache_would_fit = true;              // last ensure_pose_cache decision (false: over budget -> direct renderer)
    std::vector<double> cls_heights;          // eff-table classes of the current task table
    double min_cell = 0.0;                    // smallest cell_size of the table (bounds the crossings a ray can record)
    uint8_t *h_stage[2] = {nullptr, nullptr}; // pinned staging of mgb_maze_update_tasks (double-buffered)
    uint8_t *d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    int stage_next = 0;
    int32_t *c_vbase = nullptr;
    void *pose_rec = nullptr;      // PoseRec table, built after the variant frames
    uint8_t *c_var8 = nullptr;
    void *d_bake_desc = nullptr;
    int64_t n_var_frames = 0;
    int64_t k_hist[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // poses by the number of foods their image depends on (8 = 8 or more)
    int variant_bits_used = 0;
    double cache_bytes = 0.0;
    int variant_bits = 7;          // MGB_MAZE_VARIANT_BITS: poses that depend on <= this many foods get all 2^k frames (0 = off)
    uint8_t *c_gsig = nullptr;
    HitRec *c_hits = nullptr;
    EnvDyn *dyn = nullptr;
    void *hit_scratch = nullptr;
    size_t hit_scratch_bytes = 0;
    std::vector<int4> host_poses;
    std::vector<int32_t> host_pose_index;
    int n_tasks = 0;
    int auto_reset = 0;
    bool has_task = false, has_tex = false;
    size_t smem3d = 0;
    int render_attr_set[2] = {0, 0};   // maze3d_kernel<false / true>: shared-memory opt-in raised by this handle
    int compose_ctas_per_sm = 0;       // occupancy of maze3d_compose_kernel (queried once per handle)
    int compose_persistent = -1;       // MGB_COMPOSE_PERSISTENT
    int num_sms = 0;
    int64_t launches = 0;
    uint32_t t_base = 0;
    MgbMirrors mir = {};         // mgb_maze_set_mirrors
    MgbMirrorWindow mir_win;     // mgb_maze_set_mirror_window
};

static size_t maze3d_smem_bytes(const MazeConst &c, bool fill)
{
    auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
    size_t off = 0;
    off = up(off + (size_t)(c.n_tex + 1) * c.ts * c.ts * 4, 128);
    const int nbuf = (!fill && c.pipe) ? 2 : 1;
    off = up(off + c.blob_bytes, 128);
    off = up(off + c.blob_bytes, 128);
    for (int b = 0; b < nbuf; ++b) {
        off = up(off + (size_t)c.n * c.n * 8, 128);
        off = up(off + (size_t)c.res_h * sizeof(ColRec), 128);
        off = up(off + (size_t)c.res_v * sizeof(RowRec), 128);
        if (!c.hits_in_global) off = up(off + (size_t)c.res_h * c.max_hits * sizeof(HitRec), 128);
    }
    off = up(off + (size_t)c.res_v * 8, 128);
    const size_t px = c.obs_dtype == MGB_OBS_U8 ? 3 : 12;
    off = up(off + (size_t)(kRenderThreads / 32) * (c.obs_dtype == MGB_OBS_U8 ? 2 : 1) * c.run_px * px, 128);
    off += 32 + 128 + 16;
    return off;
}

static MazeArgs maze_args(const mgb_maze *h);
static MazeArgs maze_args_impl(const mgb_maze *h);
static MazeArgs maze_args(const mgb_maze *h)
{
    MazeArgs a = maze_args_impl(h);
    a.c_vbase = h->c_vbase;
    a.pose_rec = h->pose_rec;
    a.c_var8 = h->c_var8;
    return a;
}
static MazeArgs maze_args_impl(const mgb_maze *h)
{
    MazeArgs a;
    memset(&a, 0, sizeof(a));
    a.n = h->n; a.n_pad = h->n_pad; a.env_base = h->env_base;
    a.agent = h->agent; a.life = h->life; a.eaten = h->eaten; a.env2task = h->env2task; a.blobs = h->blobs;
    a.tex = h->tex; a.coltab = h->coltab; a.efftab = h->efftab; a.fogtab = h->fogtab; a.auto_reset = h->auto_reset;
    a.cpos = h->cpos; a.cori = h->cori; a.coltab_d = h->coltab_d;
    a.poses = h->poses; a.pose_index = h->pose_index; a.c_px = h->c_px; a.c_fid = h->c_fid;
    a.c_colhits = h->c_colhits; a.c_hits = h->c_hits; a.dyn = h->dyn; a.c_rgb8 = h->c_rgb8; a.c_gsig = h->c_gsig;
    a.c_px_all = h->c_px_all; a.c_fmask = h->c_fmask;
    return a;
}

extern "C" int mgb_maze_create(mgb_maze **out, int64_t n_envs, const mgb_maze_cfg *cfg, int device,
                               int64_t env_index_base)
{
    MGB_REQUIRE(out && cfg, "null argument");
    MGB_REQUIRE(n_envs > 0, "n_envs must be positive");
    MGB_REQUIRE(cfg->kind == MGB_MAZE_2D || cfg->kind == MGB_MAZE_DISCRETE_3D || cfg->kind == MGB_MAZE_CONTINUOUS_3D,
                "invalid kind");
    MGB_REQUIRE(cfg->task_type == MGB_MAZE_SURVIVAL || cfg->task_type == MGB_MAZE_ESCAPE, "invalid task_type");
    MGB_REQUIRE(cfg->n_cells >= 3 && cfg->n_cells <= kMaxN, "n_cells must be in [3, 31]");
    MGB_REQUIRE(cfg->max_steps > 0, "max_steps must be positive");
    if (cfg->kind == MGB_MAZE_2D) MGB_REQUIRE(cfg->view_grid >= 0 && cfg->view_grid <= 8, "view_grid out of range");
    if (cfg->kind != MGB_MAZE_2D) {
        MGB_REQUIRE(cfg->res_h > 0 && cfg->res_h <= 1024 && cfg->res_v > 0 && cfg->res_v <= 1024, "resolution out of range");
        MGB_REQUIRE(cfg->obs_dtype == MGB_OBS_U8 || cfg->obs_dtype == MGB_OBS_I32 || cfg->obs_dtype == MGB_OBS_F32,
                    "invalid obs_dtype");
        MGB_REQUIRE(cfg->max_vision > 0 && cfg->l_focal > 0 && cfg->text_size > 0 && cfg->fov > 0, "invalid optics");
    }
    int ndev = 0;
    MGB_CUDA(cudaGetDeviceCount(&ndev));
    MGB_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    MgbDeviceGuard guard(device);
    mgb_maze *h = new (std::nothrow) mgb_maze();
    MGB_REQUIRE(h, "out of host memory");
    h->device = device; h->n = n_envs; h->n_pad = (n_envs + 127) / 128 * 128; h->env_base = env_index_base;
    h->cfg = *cfg;
    MazeConst &c = h->c;
    memset(&c, 0, sizeof(c));
    c.kind = cfg->kind; c.task_type = cfg->task_type; c.n = cfg->n_cells; c.max_steps = cfg->max_steps;
    c.view_grid = cfg->view_grid; c.res_h = cfg->res_h; c.res_v = cfg->res_v; c.obs_dtype = cfg->obs_dtype;
    c.max_vision = cfg->max_vision; c.l_focal = cfg->l_focal; c.text_size = cfg->text_size;
    {
        int ex = 0;
        c.text_pow2 = (cfg->text_size > 0 && frexp(cfg->text_size, &ex) == 0.5) ? 1 : 0;
        c.inv_text = cfg->text_size > 0 ? 1.0 / cfg->text_size : 0.0;
    }
    cudaDeviceProp prop;
    MGB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    if (const char *ev = getenv("MGB_MAZE_FUSED_STEP")) h->fused_step = atoi(ev) != 0;
    if (const char *ev = getenv("MGB_MAZE_PDL")) h->step_pdl = atoi(ev) != 0;
    if (const char *ev = getenv("MGB_MAZE_RENDER_PIPE")) h->render_pipe = atoi(ev) != 0;
    if (const char *ev = getenv("MGB_MAZE_VARIANT_BITS")) {
        h->variant_bits = atoi(ev);
        if (h->variant_bits < 0) h->variant_bits = 0;
        if (h->variant_bits > kVariantBitsMax) h->variant_bits = kVariantBitsMax;
    }
    if (const char *ev = getenv("MGB_MAZE_CACHE")) h->cache_enabled = atoi(ev) != 0;
    if (const char *ev = getenv("MGB_MAZE_CACHE_GB")) h->cache_budget_gb = atof(ev);
    MGB_CUDA(cudaMalloc(&h->agent, sizeof(int4) * h->n_pad));
    MGB_CUDA(cudaMalloc(&h->life, sizeof(double) * h->n_pad));
    MGB_CUDA(cudaMalloc(&h->env2task, sizeof(int32_t) * h->n_pad));
    MGB_CUDA(cudaMemset(h->agent, 0, sizeof(int4) * h->n_pad));
    MGB_CUDA(cudaMemset(h->life, 0, sizeof(double) * h->n_pad));
    if (cfg->kind != MGB_MAZE_2D) {
        // screen geometry and the per-heading column tables, exactly as maze_view computes them
        // (ray_caster_utils.py:68-92): float64 arithmetic, float32 sin/cos of the float32 heading, float32 tables.
        const int H = cfg->res_h, V = cfg->res_v;
        volatile double half_h = tan(cfg->fov / 2) * cfg->l_focal;
        volatile double half_v = half_h * V / H;
        volatile double pixel_size = 2.0 * half_h / H;
        volatile double pixel_factor = pixel_size / cfg->l_focal;
        c.half_h = half_h; c.half_v = half_v; c.pixel_size = pixel_size;
        c.lb_sx = 0.10 * V; c.lb_sy = 0.10 * V; c.lb_w = 0.05 * H; c.lb_l = 0.80 * V;   // maze_discrete_3d.py:42-45
        std::vector<float> tab((size_t)4 * 3 * H);
        std::vector<double> tab_d((size_t)2 * H);      // heading-independent cos_hp / sin_hp in float64
        const float choice[4] = {0.0f, 0.5f, 1.0f, 1.5f};
        for (int k = 0; k < 4; ++k) {
            volatile float ori = choice[k] * (float)3.1415926;          // maze_discrete_3d.py:46
            volatile float s_ori = sinf(ori), c_ori = cosf(ori);
            volatile double tan_hp = (-0.5 - (double)H / 2) * pixel_factor;
            for (int d = 0; d < H; ++d) {
                tan_hp = tan_hp + pixel_factor;
                volatile double t2 = tan_hp * tan_hp;
                volatile double cos_hp = sqrt(1.0 / (1.0 + t2));
                volatile double sin_hp = tan_hp * cos_hp;
                volatile double a1 = sin_hp * (double)c_ori, a2 = cos_hp * (double)s_ori;
                volatile double b1 = cos_hp * (double)c_ori, b2 = sin_hp * (double)s_ori;
                if (k == 0) { tab_d[d] = cos_hp; tab_d[H + d] = sin_hp; }
                tab[((size_t)k * 3 + 0) * H + d] = (float)cos_hp;
                tab[((size_t)k * 3 + 1) * H + d] = (float)(b1 - b2);
                tab[((size_t)k * 3 + 2) * H + d] = (float)(a1 + a2);
            }
        }
        MGB_CUDA(cudaMalloc(&h->coltab, tab.size() * sizeof(float)));
        MGB_CUDA(cudaMemcpy(h->coltab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
        if (cfg->kind == MGB_MAZE_CONTINUOUS_3D) {
            MGB_CUDA(cudaMalloc(&h->coltab_d, tab_d.size() * sizeof(double)));
            MGB_CUDA(cudaMemcpy(h->coltab_d, tab_d.data(), tab_d.size() * sizeof(double), cudaMemcpyHostToDevice));
            MGB_CUDA(cudaMalloc(&h->cpos, sizeof(float2) * h->n_pad));
            MGB_CUDA(cudaMalloc(&h->cori, sizeof(double) * h->n_pad));
            MGB_CUDA(cudaMemset(h->cpos, 0, sizeof(float2) * h->n_pad));
            MGB_CUDA(cudaMemset(h->cori, 0, sizeof(double) * h->n_pad));
        }
    }
    *out = h;
    return MGB_OK;
}

extern "C" void mgb_maze_destroy(mgb_maze *h)
{
    if (!h) return;
    MgbDeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->agent); cudaFree(h->life); cudaFree(h->eaten); cudaFree(h->env2task); cudaFree(h->blobs);
    cudaFree(h->tex); cudaFree(h->coltab); cudaFree(h->efftab); cudaFree(h->fogtab); cudaFree(h->cpos); cudaFree(h->cori); cudaFree(h->coltab_d);
    cudaFree(h->poses); cudaFree(h->pose_index); cudaFree(h->c_px); cudaFree(h->c_fid); cudaFree(h->c_colhits);
    cudaFree(h->c_hits); cudaFree(h->dyn); cudaFree(h->c_rgb8); cudaFree(h->c_gsig); cudaFree(h->hit_scratch);
    cudaFree(h->c_px_all); cudaFree(h->c_fmask);
    cudaFree(h->c_vbase); cudaFree(h->c_var8); cudaFree(h->d_bake_desc); cudaFree(h->task_flags); cudaFree(h->task_epoch);
    cudaFree(h->pose_rec);
    for (int i = 0; i < 2; ++i) {
        cudaFreeHost(h->h_stage[i]); cudaFree(h->d_stage[i]);
        if (h->stage_done[i]) cudaEventDestroy(h->stage_done[i]);
    }
    delete h;
}

extern "C" int64_t mgb_maze_obs_bytes_per_env(const mgb_maze *h)
{
    if (!h) return MGB_ERR_ARG;
    if (h->c.kind == MGB_MAZE_2D) { const int w = 2 * h->c.view_grid + 1; return (int64_t)w * w * 4; }
    return (int64_t)h->c.res_h * h->c.res_v * 3 * (h->c.obs_dtype == MGB_OBS_U8 ? 1 : 4);
}

extern "C" int64_t mgb_maze_launch_count(const mgb_maze *h) { return h ? h->launches : MGB_ERR_ARG; }

extern "C" int mgb_maze_set_options(mgb_maze *h, int auto_reset)
{
    MGB_REQUIRE(h, "null handle");
    h->auto_reset = auto_reset ? 1 : 0;
    return MGB_OK;
}

extern "C" int mgb_maze_cache_info(const mgb_maze *h, int64_t out[16])
{
    MGB_REQUIRE(h && out, "null argument");
    for (int i = 0; i < 16; ++i) out[i] = 0;
    out[0] = h->cache_ready ? h->n_poses : 0;
    out[1] = h->cache_ready ? h->n_var_frames : 0;
    out[2] = h->variant_bits_used;
    out[3] = (int64_t)h->cache_bytes;
    for (int k = 0; k < 9; ++k) out[4 + k] = h->k_hist[k];
    out[13] = h->cache_ready ? 1 : 0;
    return MGB_OK;
}

extern "C" int mgb_maze_set_cache(mgb_maze *h, int enabled)
{
    MGB_REQUIRE(h, "null handle");
    if ((h->cache_enabled != 0) != (enabled != 0)) {
        h->cache_enabled = enabled ? 1 : 0;
        h->cache_dirty = true;            // rebuilt (or dropped) at the next reset / step
    }
    return MGB_OK;
}

extern "C" int mgb_maze_set_textures(mgb_maze *h, const uint8_t *grounds_host, int32_t n_tex, const uint8_t *ceil_host,
                                     int32_t tex_size)
{
    MGB_REQUIRE(h && grounds_host && ceil_host, "null argument");
    MGB_REQUIRE(h->c.kind != MGB_MAZE_2D, "textures only apply to the 3-D mazes");
    MGB_REQUIRE(n_tex >= 1 && n_tex <= 15, "n_tex must be in [1, 15]");
    MGB_REQUIRE(tex_size >= 1 && tex_size <= 128, "tex_size must be in [1, 128]");
    MgbDeviceGuard guard(h->device);
    MGB_CUDA(cudaDeviceSynchronize());
    const size_t px = (size_t)tex_size * tex_size;
    std::vector<uint32_t> packed((size_t)(n_tex + 1) * px);
    for (size_t i = 0; i < (size_t)n_tex * px; ++i)
        packed[i] = (uint32_t)grounds_host[3 * i] | ((uint32_t)grounds_host[3 * i + 1] << 8) |
                    ((uint32_t)grounds_host[3 * i + 2] << 16);
    for (size_t i = 0; i < px; ++i)
        packed[(size_t)n_tex * px + i] = (uint32_t)ceil_host[3 * i] | ((uint32_t)ceil_host[3 * i + 1] << 8) |
                                         ((uint32_t)ceil_host[3 * i + 2] << 16);
    cudaFree(h->tex); h->tex = nullptr;
    MGB_CUDA(cudaMalloc(&h->tex, packed.size() * 4));
    MGB_CUDA(cudaMemcpy(h->tex, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    h->c.n_tex = n_tex; h->c.ts = tex_size;
    h->has_tex = true;
    h->cache_dirty = true;
    return MGB_OK;
}

// One task of the table as the kernels read it: header, walls, textures, food slot of every cell, food values / intervals.
// allow_new_class: set_task may open a new (agent_height, wall_height) class (it builds the eff tables afterwards); a
// partial update may only join an existing class, otherwise the task's pixels compute eff themselves (cls = -1).
static void fill_task_blob(const MazeConst &c, uint8_t *b, const int8_t *walls, const int8_t *texts, const double *food,
                           const int32_t *interval, const mgb_maze_task_scalars &s, std::vector<double> &cls_heights,
                           bool allow_new_class)
{
    const int nn = c.n * c.n;
    TaskHdr hd;
    memset(&hd, 0, sizeof(hd));
    hd.start[0] = s.start[0]; hd.start[1] = s.start[1]; hd.goal[0] = s.goal[0]; hd.goal[1] = s.goal[1];
    hd.cell_size = s.cell_size; hd.wall_height = s.wall_height; hd.agent_height = s.agent_height;
    hd.initial_life = s.initial_life; hd.max_life = s.max_life; hd.step_reward = s.step_reward;
    hd.goal_reward = s.goal_reward;
    {
        int ex = 0;
        const double t2c = c.text_size / s.cell_size;
        hd.cell_pow2 = frexp(s.cell_size, &ex) == 0.5 ? 1 : 0;
        hd.t2c_pow2 = frexp(t2c, &ex) == 0.5 ? 1 : 0;
        hd.inv_cell = 1.0 / s.cell_size;
        hd.inv_t2c = 1.0 / t2c;
        hd.cls = -1;
        for (size_t k = 0; k < cls_heights.size() / 2; ++k)
            if (cls_heights[2 * k] == s.agent_height && cls_heights[2 * k + 1] == s.wall_height) hd.cls = (int)k;
        if (hd.cls < 0 && allow_new_class && cls_heights.size() / 2 < 8) {
            hd.cls = (int)(cls_heights.size() / 2);
            cls_heights.push_back(s.agent_height);
            cls_heights.push_back(s.wall_height);
        }
    }
    int cnt = 0;
    int8_t *fidx = reinterpret_cast<int8_t *>(b + c.off_fidx);
    double *fval = reinterpret_cast<double *>(b + c.off_fval);
    int32_t *fint = reinterpret_cast<int32_t *>(b + c.off_fint);
    for (int k = 0; k < nn; ++k) {
        b[c.off_walls + k] = (uint8_t)walls[k];
        b[c.off_texts + k] = (uint8_t)texts[k];
        const double fv = food[k];
        if (fv > 0.0) {
            fidx[k] = (int8_t)cnt; fval[cnt] = fv; fint[cnt] = interval[k];
            ++cnt;
        } else fidx[k] = -1;
    }
    hd.n_food = cnt;
    memcpy(b, &hd, sizeof(hd));
}

extern "C" int mgb_maze_set_task(mgb_maze *h, int32_t n_tasks, const int8_t *walls_host, const int8_t *texts_host,
                                 const double *food_rewards_host, const int32_t *food_interval_host,
                                 const mgb_maze_task_scalars *scalars_host, const int32_t *env2task_host)
{
    MgbRange nvtx_range("mgb_maze_set_task");
    MGB_REQUIRE(h && walls_host && texts_host && food_rewards_host && food_interval_host && scalars_host &&
                    env2task_host, "null argument");
    MGB_REQUIRE(n_tasks > 0, "n_tasks must be positive");
    MgbDeviceGuard guard(h->device);
    MGB_CUDA(cudaDeviceSynchronize());
    MazeConst &c = h->c;
    const int n = c.n, nn = n * n;
    // food slots: every cell with a positive reward (the ceiling tint of ray_caster_utils.py:150 tests "> 0")
    int f_max = 0;
    for (int t = 0; t < n_tasks; ++t) {
        int cnt = 0;
        for (int k = 0; k < nn; ++k) cnt += food_rewards_host[(size_t)t * nn + k] > 0.0 ? 1 : 0;
        f_max = cnt > f_max ? cnt : f_max;
        const mgb_maze_task_scalars &s = scalars_host[t];
        MGB_REQUIRE(s.start[0] >= 0 && s.start[0] < n && s.start[1] >= 0 && s.start[1] < n, "start outside the maze");
        MGB_REQUIRE(s.goal[0] >= 0 && s.goal[0] < n && s.goal[1] >= 0 && s.goal[1] < n, "goal outside the maze");
        // maze_base.py:36
        MGB_REQUIRE(s.agent_height < s.wall_height && s.agent_height > 0, "the agent height must be > 0 and < wall height");
        MGB_REQUIRE(s.cell_size > 0, "cell_size must be positive");
    }
    MGB_REQUIRE(f_max <= 127, "at most 127 food cells per task are supported");
    for (int e = 0; e < h->n; ++e) MGB_REQUIRE(env2task_host[e] >= 0 && env2task_host[e] < n_tasks, "env2task out of range");
    {
        std::vector<uint8_t> used((size_t)n_tasks, 0);
        h->slot_per_env = true;
        for (int e = 0; e < h->n; ++e) {
            if (used[env2task_host[e]]) { h->slot_per_env = false; break; }
            used[env2task_host[e]] = 1;
        }
    }
    c.f_max = f_max;
    // A ray is followed for max_vision at most, i.e. through <= 2 * max_vision / cell_size + 2 cells (one per DDA step
    // plus the start cell): that bounds the transparent crossings a column can record (ray_caster_utils.py:24-61).
    double min_cell = scalars_host[0].cell_size;
    for (int t = 1; t < n_tasks; ++t) min_cell = scalars_host[t].cell_size < min_cell ? scalars_host[t].cell_size : min_cell;
    h->min_cell = min_cell;
    const int geo = (int)ceil(2.0 * c.max_vision / min_cell) + 3;
    int mh = f_max < geo ? f_max : geo;
    if (mh < 1) mh = 1;
    if (c.task_type == MGB_MAZE_ESCAPE) mh = 2;
    c.max_hits = mh > kMaxHitsCap ? kMaxHitsCap : mh;
    // a warp run = whole columns, about 768 B of output (u8: 256 px, i32: 64 px), never less than one column
    c.run_px = c.obs_dtype == MGB_OBS_U8 ? 256 : 64;
    if (c.kind != MGB_MAZE_2D) {
        if (c.run_px < c.res_v) c.run_px = c.res_v;
        c.run_px = c.run_px / c.res_v * c.res_v;
    }
    // blob layout
    size_t off = sizeof(TaskHdr);
    c.off_walls = (int)off; off += nn;
    c.off_texts = (int)off; off += nn;
    c.off_fidx = (int)off;  off += nn;
    off = (off + 7) / 8 * 8;
    c.off_fval = (int)off;  off += (size_t)(f_max > 0 ? f_max : 1) * 8;
    c.off_fint = (int)off;  off += (size_t)(f_max > 0 ? f_max : 1) * 4;
    c.blob_bytes = (int)((off + 15) / 16 * 16);
    std::vector<uint8_t> blobs((size_t)n_tasks * c.blob_bytes, 0);
    std::vector<double> &cls_heights = h->cls_heights;   // distinct (agent_height, wall_height) pairs, at most 8 get an eff table
    cls_heights.clear();
    for (int t = 0; t < n_tasks; ++t)
        fill_task_blob(c, blobs.data() + (size_t)t * c.blob_bytes, walls_host + (size_t)t * nn, texts_host + (size_t)t * nn,
                       food_rewards_host + (size_t)t * nn, food_interval_host + (size_t)t * nn, scalars_host[t], cls_heights,
                       true);
    if (c.kind != MGB_MAZE_2D) {
        for (int t = 0; t < n_tasks; ++t)
            for (int k = 0; k < nn; ++k) {
                const int id = texts_host[(size_t)t * nn + k];
                MGB_REQUIRE(id >= 0 && (!h->has_tex || id < c.n_tex), "cell_texts refers to a texture that is not loaded");
            }
    }
    cudaFree(h->blobs); h->blobs = nullptr;
    cudaFree(h->eaten); h->eaten = nullptr;
    MGB_CUDA(cudaMalloc(&h->blobs, blobs.size()));
    MGB_CUDA(cudaMemcpy(h->blobs, blobs.data(), blobs.size(), cudaMemcpyHostToDevice));
    MGB_CUDA(cudaMalloc(&h->eaten, sizeof(int32_t) * (size_t)(f_max > 0 ? f_max : 1) * h->n_pad));
    MGB_CUDA(cudaMemcpy(h->env2task, env2task_host, sizeof(int32_t) * h->n, cudaMemcpyHostToDevice));
    h->n_tasks = n_tasks;
    h->has_task = true;
    cudaFree(h->task_flags); h->task_flags = nullptr; h->task_flags_n = 0;
    h->cache_would_fit = true;
    // pose list of the cache: every free cell (the agent can never stand inside a wall, maze_discrete_3d.py:63-65) x 4
    h->host_poses.clear();
    h->host_pose_index.assign((size_t)n_tasks * nn * 4, -1);
    h->cache_dirty = true;
    h->cache_ready = false;
    if (c.kind == MGB_MAZE_DISCRETE_3D) {
        for (int t = 0; t < n_tasks; ++t) {
            const mgb_maze_task_scalars &sc = scalars_host[t];
            for (int k = 0; k < nn; ++k) {
                const bool is_start = (k == sc.start[0] * n + sc.start[1]);
                if (walls_host[(size_t)t * nn + k] != 0 && !is_start) continue;
                for (int o = 0; o < 4; ++o) {
                    h->host_pose_index[((size_t)t * nn + k) * 4 + o] = (int32_t)h->host_poses.size();
                    h->host_poses.push_back(make_int4(t, k / n, k % n, o));
                }
            }
        }
    }
    cudaFree(h->efftab); h->efftab = nullptr;
    cudaFree(h->fogtab); h->fogtab = nullptr;
    c.n_cls = 0;
    if (c.kind != MGB_MAZE_2D && !cls_heights.empty()) {
        const int n_cls = (int)(cls_heights.size() / 2);
        const size_t cells = (size_t)n_cls * c.res_h * c.res_v;
        double *d_heights = nullptr;
        MGB_CUDA(cudaMalloc(&h->efftab, cells * sizeof(double)));
        MGB_CUDA(cudaMalloc(&h->fogtab, cells * sizeof(double)));
        MGB_CUDA(cudaMalloc(&d_heights, cls_heights.size() * sizeof(double)));
        MGB_CUDA(cudaMemcpy(d_heights, cls_heights.data(), cls_heights.size() * sizeof(double), cudaMemcpyHostToDevice));
        maze_efftab_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(c, h->coltab, d_heights, n_cls, h->efftab, h->fogtab);
        MGB_CUDA(cudaDeviceSynchronize());
        cudaFree(d_heights);
        c.n_cls = n_cls;
        h->launches += 1;
    }
    // set_task leaves the env in "need reset" state (maze_env.py:44-50): initialise it so a stray step is harmless
    MazeArgs a = maze_args(h);
    maze_reset_kernel<<<(unsigned)((h->n + 255) / 256), 256>>>(c, a);
    MGB_CUDA(cudaDeviceSynchronize());
    h->launches += 1;
    return MGB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device-side task sampler: per-episode task resampling without the host.  Measured motivation
// (bench.py --workload maze3d, task_churn): 1024 envs finish ~26 000 episodes/s on the direct renderer, the host samplers
// deliver 700 (reference random streams) to 4 200 (numpy RandomState) tasks/s.  This kernel draws a fresh maze per finished
// env -- one thread per env, Philox keyed by (seed, global env index, resample count) -- from the same distribution family
// as MazeTaskSampler(rng=...) (metagym_b200/metamaze.py; the reference's maze_task.py:41-190 draws from Python's and
// numpy's global MT19937 streams, which a device cannot replay: parity here is distributional and structural, asserted
// in tests/test_maze_gpu.py: rooms on odd coordinates, border walls, a spanning tree of the room lattice by randomised
// Kruskal, loops knocked out down to crowd_ratio, textures 1..n_texts-1 on walls, start/goal rooms > 0.45 n apart, food
// values clip(U * food_reward, 0.1, food_reward) thinned by 0.9 per round until their sum is <= (n-1)^2 food_density).
// ---------------------------------------------------------------------------------------------------------------
struct SamplerCfg {          // mgb_maze_sampler_cfg + derived fields
    int allow_loops, n_texts, food_interval, cls;
    double cell_size, wall_height, agent_height, step_reward, goal_reward, food_reward, initial_life, max_life, food_density,
        crowd_ratio;
};
struct PhiloxStream {
    uint2 key;
    uint4 ctr;
    uint4 buf;
    int have;
    __device__ uint32_t next()
    {
        if (have == 0) { buf = mgb_philox4x32_10(ctr, key); ctr.z += 1u; have = 4; }
        const uint32_t v = have == 4 ? buf.x : (have == 3 ? buf.y : (have == 2 ? buf.z : buf.w));
        --have;
        return v;
    }
    __device__ double uniform() { return (double)(next() >> 8) * (1.0 / 16777216.0); }
    __device__ int below(int n) { return (int)(((uint64_t)next() * (uint64_t)n) >> 32); }     // uniform in [0, n)
};
#define MGB_STREAM_SAMPLER 0x300u

constexpr int kSamplerWarps = 4;
// One WARP per env.  Lane 0 carves the maze (randomised Kruskal + loop knock-out are sequential by nature; their working set
// lives in shared memory); textures, food values and the food thinning rounds are counter-based per cell -- Philox keyed by
// (seed; global env, resample count, cell, purpose) -- so the 32 lanes take cells side by side.  A task costs ~50 us of one
// warp instead of ~1.7 ms of one thread (the thinning alone was 20 rounds x n^2 serial draws).
__global__ void __launch_bounds__(32 * kSamplerWarps) maze_sample_tasks_kernel(const __grid_constant__ MazeConst c,
                                                                              const __grid_constant__ MazeArgs a,
                                                                              uint8_t *blobs, const uint8_t *mask, uint32_t *epoch,
                                                                              const __grid_constant__ SamplerCfg sc, uint64_t seed)
{
    __shared__ uint8_t s_walls[kSamplerWarps][kMaxN * kMaxN];      // bit 0 wall, bit 1 food alive
    __shared__ uint8_t s_parent[kSamplerWarps][((kMaxN - 1) / 2) * ((kMaxN - 1) / 2)];
    __shared__ uint16_t s_order[kSamplerWarps][kMaxN * kMaxN];
    __shared__ uint32_t s_val[kSamplerWarps][kMaxN * kMaxN];       // 24-bit draw behind every cell's food value
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e = (int64_t)blockIdx.x * kSamplerWarps + w;
    if (e >= a.n) return;                                          // warp-uniform exits: no block-wide barrier below
    if (mask && !mask[e]) return;
    uint8_t *walls = s_walls[w];
    uint8_t *parent = s_parent[w];
    uint16_t *order = s_order[w];
    uint32_t *val = s_val[w];
    const int n = c.n, nn = n * n, m = (n - 1) / 2;
    const uint32_t ep = epoch[e] + 1u;
    __syncwarp();
    if (lane == 0) epoch[e] = ep;
    const int64_t genv = a.env_base + e;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t hi = (uint32_t)((uint64_t)genv >> 32);
    // per-cell draws: ctr = (env, resample count, cell, purpose)
    auto cell_draw = [&](int k, uint32_t purpose) { return mgb_philox4x32_10(make_uint4((uint32_t)genv, ep, (uint32_t)k, purpose + hi), key); };
    const uint32_t P_TEX = MGB_STREAM_SAMPLER + 0x10u, P_VAL = MGB_STREAM_SAMPLER + 0x20u, P_KEEP = MGB_STREAM_SAMPLER + 0x1000u;
    for (int k = lane; k < nn; k += 32) {
        const int i = k / n, j = k - i * n;
        walls[k] = ((i & 1) && (j & 1)) ? 0 : 1;
    }
    for (int k = lane; k < m * m; k += 32) parent[k] = (uint8_t)k;
    __syncwarp();
    int sx = 1, sy = 1, gx = n - 2, gy = n - 2;
    if (lane == 0) {
        PhiloxStream rng;                                            // the sequential stream of the carving steps
        rng.key = key;
        rng.ctr = make_uint4((uint32_t)genv, ep, 0u, MGB_STREAM_SAMPLER + hi);
        rng.have = 0;
        auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
        // ---- random spanning tree of the room lattice (Kruskal over shuffled edges); edge id = 2 * room + dir
        int ne = 0;
        for (int ra = 0; ra < m; ++ra)
            for (int rb = 0; rb < m; ++rb) {
                if (ra + 1 < m) order[ne++] = (uint16_t)(2 * (ra * m + rb));
                if (rb + 1 < m) order[ne++] = (uint16_t)(2 * (ra * m + rb) + 1);
            }
        for (int k = ne - 1; k > 0; --k) { const int j = rng.below(k + 1); const uint16_t t = order[k]; order[k] = order[j]; order[j] = t; }
        for (int k = 0; k < ne; ++k) {
            const int room = order[k] >> 1, dir = order[k] & 1, ra = room / m, rb = room % m;
            const int other = dir == 0 ? (ra + 1) * m + rb : ra * m + rb + 1;
            const int x = find(room), y = find(other);
            if (x != y) {
                parent[x] = (uint8_t)y;
                if (dir == 0) walls[(2 * ra + 2) * n + 2 * rb + 1] = 0; else walls[(2 * ra + 1) * n + 2 * rb + 2] = 0;
            }
        }
        // ---- loops: knock interior walls out (in random order, only next to a free cell) down to crowd_ratio
        if (sc.allow_loops) {
            int standing = 0, nc = 0;
            for (int i = 1; i < n - 1; ++i)
                for (int j = 1; j < n - 1; ++j)
                    if (walls[i * n + j]) { ++standing; order[nc++] = (uint16_t)(i * n + j); }
            const double budget = (double)((n - 2) * (n - 2)) * sc.crowd_ratio;
            for (int k = nc - 1; k > 0; --k) { const int j = rng.below(k + 1); const uint16_t t = order[k]; order[k] = order[j]; order[j] = t; }
            for (int k = 0; k < nc && (double)standing > budget; ++k) {
                const int cell = order[k];
                if (!walls[cell - n] || !walls[cell + n] || !walls[cell - 1] || !walls[cell + 1]) { walls[cell] = 0; --standing; }
            }
        }
        // ---- start / goal (maze_task.py:153-160: rooms, goal far enough from the start)
        sx = rng.below(m) * 2 + 1; sy = rng.below(m) * 2 + 1;
        for (int t = 0; t < m * m; ++t) {
            const int ex = rng.below(m) * 2 + 1, ey = rng.below(m) * 2 + 1;
            const double dx = ex - sx, dy = ey - sy;
            if (sqrt(dx * dx + dy * dy) > 0.45 * n) { gx = ex; gy = ey; break; }
        }
    }
    __syncwarp();
    // ---- textures + food values, one cell per lane and pass
    uint8_t *b = blobs + (size_t)a.env2task[e] * c.blob_bytes;
    auto value_of = [&](uint32_t u24) {                               // clip(food_reward * U, 0.10, food_reward)
        const double v = (double)u24 * (1.0 / 16777216.0) * sc.food_reward;
        return v < 0.10 ? 0.10 : (v > sc.food_reward ? sc.food_reward : v);
    };
    double total = 0.0;
    int alive = 0;
    for (int k = lane; k < nn; k += 32) {
        const uint4 r = cell_draw(k, P_TEX);
        const int tx = 1 + (int)(((uint64_t)r.x * (uint64_t)(sc.n_texts - 1)) >> 32);      // randint(1, n_texts)
        const bool wall = walls[k] & 1;
        b[c.off_walls + k] = wall ? 1 : 0;
        b[c.off_texts + k] = wall ? (uint8_t)tx : 0;
        const uint32_t u24 = cell_draw(k, P_VAL).x >> 8;
        val[k] = u24;
        if (!wall) { walls[k] |= 2; total += value_of(u24); ++alive; }
    }
    auto warp_sum = [&](double &t, int &cnt) {                        // fixed butterfly: the same sum on every lane
        for (int o = 16; o > 0; o >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    };
    warp_sum(total, alive);
    // ---- thinning: food *= (rand < 0.90) per round until the total value and the slot count fit (maze_task.py:176-180)
    const double expected = (double)((n - 1) * (n - 1)) * sc.food_density;
    for (uint32_t round = 0; total > expected || alive > c.f_max; ++round) {
        total = 0.0; alive = 0;
        for (int k = lane; k < nn; k += 32) {
            if (!(walls[k] & 2)) continue;
            const uint4 r = cell_draw(k, P_KEEP + (round >> 2));
            const uint32_t x = (round & 3u) == 0 ? r.x : ((round & 3u) == 1 ? r.y : ((round & 3u) == 2 ? r.z : r.w));
            if ((double)(x >> 8) * (1.0 / 16777216.0) < 0.90) { total += value_of(val[k]); ++alive; }
            else walls[k] &= ~2;
        }
        warp_sum(total, alive);
    }
    __syncwarp();
    int8_t *fidx = reinterpret_cast<int8_t *>(b + c.off_fidx);
    double *fval = reinterpret_cast<double *>(b + c.off_fval);
    int32_t *fint = reinterpret_cast<int32_t *>(b + c.off_fint);
    if (lane != 0) return;
    {
        int cnt = 0;
        for (int k = 0; k < nn; ++k) {                                 // food slots numbered in cell order
            if (walls[k] & 2) { fidx[k] = (int8_t)cnt; fval[cnt] = value_of(val[k]); fint[cnt] = sc.food_interval; ++cnt; }
            else fidx[k] = -1;
        }
        TaskHdr hd;
        hd.start[0] = sx; hd.start[1] = sy; hd.goal[0] = gx; hd.goal[1] = gy;
        hd.cell_size = sc.cell_size; hd.wall_height = sc.wall_height; hd.agent_height = sc.agent_height;
        hd.initial_life = sc.initial_life; hd.max_life = sc.max_life; hd.step_reward = sc.step_reward;
        hd.goal_reward = sc.goal_reward;
        hd.n_food = cnt; hd.cls = sc.cls;
        int ex = 0;
        const double t2c = c.text_size / sc.cell_size;
        hd.cell_pow2 = frexp(sc.cell_size, &ex) == 0.5 ? 1 : 0;
        hd.t2c_pow2 = frexp(t2c, &ex) == 0.5 ? 1 : 0;
        hd.inv_cell = 1.0 / sc.cell_size;
        hd.inv_t2c = 1.0 / t2c;
        *reinterpret_cast<TaskHdr *>(b) = hd;
    }
    // ---- the env starts an episode on its new task (set_task + reset of that env, maze_env.py:44-57)
    Env s;
    env_reset(c, b, a.eaten + e, a.n_pad, s);
    a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
    a.life[e] = s.life;
    if (c.kind == MGB_MAZE_CONTINUOUS_3D) {
        a.cpos[e] = make_float2((float)(s.gx * sc.cell_size + 0.5 * sc.cell_size), (float)(s.gy * sc.cell_size + 0.5 * sc.cell_size));
        a.cori[e] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Partial, stream-ordered task replacement (per-episode task resampling, SURVEY.md 8f row 3 / maze_base.py:19-38)
// ---------------------------------------------------------------------------------------------------------------
// staging layout: [count] int32 table slots | pad to 16 | [count] blobs
__global__ void maze_scatter_tasks_kernel(const __grid_constant__ MazeConst c, uint8_t *blobs, const uint8_t *stage, int count,
                                          uint8_t *task_flags)
{
    const int32_t *slots = reinterpret_cast<const int32_t *>(stage);
    const uint4 *src = reinterpret_cast<const uint4 *>(stage + (((size_t)count * 4 + 15) / 16) * 16) + (size_t)blockIdx.x * (c.blob_bytes / 16);
    uint4 *dst = reinterpret_cast<uint4 *>(blobs + (size_t)slots[blockIdx.x] * c.blob_bytes);
    for (int i = threadIdx.x; i < c.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) task_flags[slots[blockIdx.x]] = 1;
}
// envs whose task was replaced start a new episode on it (set_task leaves an env at its start state, maze_env.py:44-50)
__global__ void maze_reset_flagged_kernel(const __grid_constant__ MazeConst c, const __grid_constant__ MazeArgs a,
                                          const uint8_t *task_flags)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    const int task = a.env2task[e];
    if (!task_flags[task]) return;
    const uint8_t *blob = a.blobs + (int64_t)task * c.blob_bytes;
    Env s;
    env_reset(c, blob, a.eaten + e, a.n_pad, s);
    a.agent[e] = make_int4(s.gx, s.gy, s.ori, s.steps);
    a.life[e] = s.life;
    if (c.kind == MGB_MAZE_CONTINUOUS_3D) {             // get_cell_center(start), heading 0 (maze_base.py:41,50)
        const TaskHdr *th = blob_hdr(blob);
        a.cpos[e] = make_float2((float)(s.gx * th->cell_size + 0.5 * th->cell_size),
                                (float)(s.gy * th->cell_size + 0.5 * th->cell_size));
        a.cori[e] = 0.0;
    }
}
__global__ void maze_clear_flags_kernel(uint8_t *task_flags, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_flags[i] = 0;
}

extern "C" int mgb_maze_update_tasks(mgb_maze *h, int32_t count, const int32_t *task_slots_host, const int8_t *walls_host,
                                     const int8_t *texts_host, const double *food_rewards_host,
                                     const int32_t *food_interval_host, const mgb_maze_task_scalars *scalars_host,
                                     void *stream)
{
    MgbRange nvtx_range("mgb_maze_update_tasks");
    MGB_REQUIRE(h && task_slots_host && walls_host && texts_host && food_rewards_host && food_interval_host && scalars_host,
                "null argument");
    MGB_REQUIRE(count > 0, "count must be positive");
    MGB_REQUIRE(h->has_task, "call mgb_maze_set_task first (it sizes the task table)");
    MgbDeviceGuard guard(h->device);
    MazeConst &c = h->c;
    MGB_REQUIRE(!(c.kind == MGB_MAZE_DISCRETE_3D && h->cache_enabled && !h->host_poses.empty() && h->cache_would_fit),
                "partial task updates need the direct renderer: create the env with the pose cache off (MGB_MAZE_CACHE=0 / "
                "cache=False) -- the cache memoises whole task tables");
    const int n = c.n, nn = n * n;
    for (int t = 0; t < count; ++t) {
        MGB_REQUIRE(task_slots_host[t] >= 0 && task_slots_host[t] < h->n_tasks, "task slot out of range");
        int cnt = 0;
        for (int k = 0; k < nn; ++k) {
            cnt += food_rewards_host[(size_t)t * nn + k] > 0.0 ? 1 : 0;
            const int id = texts_host[(size_t)t * nn + k];
            MGB_REQUIRE(c.kind == MGB_MAZE_2D || (id >= 0 && (!h->has_tex || id < c.n_tex)), "cell_texts refers to a texture that is not loaded");
        }
        MGB_REQUIRE(cnt <= c.f_max, "a replacement task may not have more food cells than the largest task of set_task");
        const mgb_maze_task_scalars &s = scalars_host[t];
        MGB_REQUIRE(s.start[0] >= 0 && s.start[0] < n && s.start[1] >= 0 && s.start[1] < n, "start outside the maze");
        MGB_REQUIRE(s.goal[0] >= 0 && s.goal[0] < n && s.goal[1] >= 0 && s.goal[1] < n, "goal outside the maze");
        MGB_REQUIRE(s.agent_height < s.wall_height && s.agent_height > 0, "the agent height must be > 0 and < wall height");
        MGB_REQUIRE(s.cell_size >= h->min_cell, "a replacement task may not have smaller cells than the table's smallest");
    }
    // pinned staging, double-buffered: the host waits only for the COPY of the call before last, never for the device
    const int sb = h->stage_next;
    h->stage_next ^= 1;
    const size_t head = (((size_t)count * 4 + 15) / 16) * 16, need = head + (size_t)count * c.blob_bytes;
    if (!h->stage_done[sb]) MGB_CUDA(cudaEventCreateWithFlags(&h->stage_done[sb], cudaEventDisableTiming));
    else MGB_CUDA(cudaEventSynchronize(h->stage_done[sb]));
    if (need > h->stage_bytes[sb]) {
        cudaFreeHost(h->h_stage[sb]); cudaFree(h->d_stage[sb]);
        h->h_stage[sb] = nullptr; h->d_stage[sb] = nullptr; h->stage_bytes[sb] = 0;
        const size_t cap = need * 2;
        MGB_CUDA(cudaMallocHost(&h->h_stage[sb], cap));
        MGB_CUDA(cudaMalloc(&h->d_stage[sb], cap));
        h->stage_bytes[sb] = cap;
    }
    if (!h->task_flags) {
        MGB_CUDA(cudaMalloc(&h->task_flags, (size_t)h->n_tasks));
        MGB_CUDA(cudaMemset(h->task_flags, 0, (size_t)h->n_tasks));
        h->task_flags_n = h->n_tasks;
    }
    uint8_t *hs = h->h_stage[sb];
    memset(hs, 0, need);
    memcpy(hs, task_slots_host, (size_t)count * 4);
    for (int t = 0; t < count; ++t)
        fill_task_blob(c, hs + head + (size_t)t * c.blob_bytes, walls_host + (size_t)t * nn, texts_host + (size_t)t * nn,
                       food_rewards_host + (size_t)t * nn, food_interval_host + (size_t)t * nn, scalars_host[t], h->cls_heights,
                       false);
    cudaStream_t st = (cudaStream_t)stream;
    MGB_CUDA(cudaMemcpyAsync(h->d_stage[sb], hs, need, cudaMemcpyHostToDevice, st));
    MGB_CUDA(cudaEventRecord(h->stage_done[sb], st));
    maze_scatter_tasks_kernel<<<(unsigned)count, 128, 0, st>>>(c, h->blobs, h->d_stage[sb], count, h->task_flags);
    MazeArgs a = maze_args(h);
    maze_reset_flagged_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, st>>>(c, a, h->task_flags);
    maze_clear_flags_kernel<<<(unsigned)((h->n_tasks + 255) / 256), 256, 0, st>>>(h->task_flags, h->n_tasks);
    MGB_CUDA(cudaGetLastError());
    h->launches += 3;
    return MGB_OK;
}

extern "C" int mgb_maze_resample_tasks(mgb_maze *h, const uint8_t *mask_dev, const mgb_maze_sampler_cfg *cfg, uint64_t seed,
                                       void *stream)
{
    MgbRange nvtx_range("mgb_maze_resample_tasks");
    MGB_REQUIRE(h && cfg, "null argument");
    MGB_REQUIRE(h->has_task, "call mgb_maze_set_task first (it sizes the task table)");
    MGB_REQUIRE(h->slot_per_env, "device resampling needs one task-table slot per env (mgb_maze_set_task with n_tasks >= n_envs "
                                 "and an injective env2task)");
    MgbDeviceGuard guard(h->device);
    MazeConst &c = h->c;
    MGB_REQUIRE(!(c.kind == MGB_MAZE_DISCRETE_3D && h->cache_enabled && !h->host_poses.empty() && h->cache_would_fit),
                "device resampling needs the direct renderer: create the env with the pose cache off (cache=False)");
    MGB_REQUIRE(c.n % 2 == 1 && c.n > 6, "Cell Numbers can only be odd, minimum 7 (maze_task.py:57-58)");
    MGB_REQUIRE(cfg->step_reward < 0, "step_reward must be < 0 (maze_task.py:59)");
    MGB_REQUIRE(cfg->agent_height < cfg->wall_height && cfg->agent_height > 0, "the agent height must be > 0 and < wall height");
    MGB_REQUIRE(cfg->cell_size >= h->min_cell, "resampled tasks may not have smaller cells than the table's smallest");
    MGB_REQUIRE(cfg->n_texts >= 2 && (c.kind == MGB_MAZE_2D || !h->has_tex || cfg->n_texts <= c.n_tex), "n_texts out of range");
    MGB_REQUIRE(cfg->food_reward > 0 && cfg->food_density >= 0 && cfg->crowd_ratio >= 0, "invalid sampler parameters");
    SamplerCfg sc;
    sc.allow_loops = cfg->allow_loops; sc.n_texts = cfg->n_texts; sc.food_interval = cfg->food_interval;
    sc.cell_size = cfg->cell_size; sc.wall_height = cfg->wall_height; sc.agent_height = cfg->agent_height;
    sc.step_reward = cfg->step_reward;
    sc.goal_reward = cfg->goal_reward > 0 ? cfg->goal_reward : -sqrt((double)c.n) * c.n * cfg->step_reward;   // maze_task.py:163-166
    sc.food_reward = cfg->food_reward; sc.initial_life = cfg->initial_life; sc.max_life = cfg->max_life;
    sc.food_density = cfg->food_density; sc.crowd_ratio = cfg->crowd_ratio;
    sc.cls = -1;
    for (size_t k = 0; k < h->cls_heights.size() / 2; ++k)
        if (h->cls_heights[2 * k] == cfg->agent_height && h->cls_heights[2 * k + 1] == cfg->wall_height) sc.cls = (int)k;
    if (!h->task_epoch) {
        MGB_CUDA(cudaMalloc(&h->task_epoch, sizeof(uint32_t) * (size_t)h->n_pad));
        MGB_CUDA(cudaMemset(h->task_epoch, 0, sizeof(uint32_t) * (size_t)h->n_pad));
    }
    MazeArgs a = maze_args(h);
    maze_sample_tasks_kernel<<<(unsigned)((h->n + kSamplerWarps - 1) / kSamplerWarps), 32 * kSamplerWarps, 0, (cudaStream_t)stream>>>(c, a, h->blobs, mask_dev, h->task_epoch,
                                                                                        sc, seed);
    MGB_CUDA(cudaGetLastError());
    h->launches += 1;
    return MGB_OK;
}

extern "C" int mgb_maze_get_tasks(mgb_maze *h, int32_t count, const int32_t *task_slots_host, int8_t *walls_host,
                                  int8_t *texts_host, double *food_rewards_host, int32_t *food_interval_host,
                                  mgb_maze_task_scalars *scalars_host)
{
    MGB_REQUIRE(h && task_slots_host && walls_host && texts_host && food_rewards_host && food_interval_host && scalars_host,
                "null argument");
    MGB_REQUIRE(h->has_task && count > 0, "no task table");
    MgbDeviceGuard guard(h->device);
    MGB_CUDA(cudaDeviceSynchronize());
    const MazeConst &c = h->c;
    const int nn = c.n * c.n;
    std::vector<uint8_t> blob((size_t)c.blob_bytes);
    for (int t = 0; t < count; ++t) {
        MGB_REQUIRE(task_slots_host[t] >= 0 && task_slots_host[t] < h->n_tasks, "task slot out of range");
        MGB_CUDA(cudaMemcpy(blob.data(), h->blobs + (size_t)task_slots_host[t] * c.blob_bytes, blob.size(), cudaMemcpyDeviceToHost));
        TaskHdr hd;
        memcpy(&hd, blob.data(), sizeof(hd));
        const int8_t *fidx = reinterpret_cast<const int8_t *>(blob.data() + c.off_fidx);
        const double *fval = reinterpret_cast<const double *>(blob.data() + c.off_fval);
        const int32_t *fint = reinterpret_cast<const int32_t *>(blob.data() + c.off_fint);
        for (int k = 0; k < nn; ++k) {
            walls_host[(size_t)t * nn + k] = (int8_t)blob[c.off_walls + k];
            texts_host[(size_t)t * nn + k] = (int8_t)blob[c.off_texts + k];
            const int f = fidx[k];
            food_rewards_host[(size_t)t * nn + k] = f >= 0 ? fval[f] : 0.0;
            food_interval_host[(size_t)t * nn + k] = f >= 0 ? fint[f] : 0;
        }
        mgb_maze_task_scalars &s = scalars_host[t];
        s.start[0] = hd.start[0]; s.start[1] = hd.start[1]; s.goal[0] = hd.goal[0]; s.goal[1] = hd.goal[1];
        s.cell_size = hd.cell_size; s.wall_height = hd.wall_height; s.agent_height = hd.agent_height;
        s.initial_life = hd.initial_life; s.max_life = hd.max_life; s.step_reward = hd.step_reward; s.goal_reward = hd.goal_reward;
    }
    return MGB_OK;
}

static int maze_ready(const mgb_maze *h)
{
    if (!h->has_task) { mgb_set_error("Must call \"set_task\" before reset"); return MGB_ERR_STATE; }   // maze_env.py:49-50
    if (h->c.kind != MGB_MAZE_2D && !h->has_tex) {
        mgb_set_error("3-D maze: call mgb_maze_set_textures first");
        return MGB_ERR_STATE;
    }
    return MGB_OK;
}

// pose_index + c_fmask + c_vbase -> one PoseRec per (task, cell, heading)
__global__ void maze_pose_rec_kernel(const int32_t *pose_index, const uint64_t *fmask, const int32_t *vbase, PoseRec *rec, int64_t count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    PoseRec r;
    r.slot = pose_index[i]; r.vbase = -1; r.fmask[0] = r.fmask[1] = 0; r.pad = 0;
    if (r.slot >= 0) {
        r.fmask[0] = fmask[(size_t)r.slot * 2]; r.fmask[1] = fmask[(size_t)r.slot * 2 + 1];
        if (vbase) r.vbase = vbase[r.slot];
    }
    rec[i] = r;
}

// Raise a kernel's dynamic shared-memory opt-in to everything the device allows next to the kernel's static shared memory.
// The value does not depend on the calling handle, so handles (and host threads) cannot undo each other's setting.
template <class F> static cudaError_t maze_allow_max_dynamic_smem(F *kernel)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

template <bool FILL>
static int launch_render(mgb_maze *h, const MazeArgs &a, unsigned grid, cudaStream_t st)
{
    MazeConst &c = h->c;
    // shared-memory plan, most wanted first: (direct renderer) two record sets with the crossing lists in shared memory,
    // two record sets with the lists in a global scratch, then one record set (the only plan of the FILL pass)
    size_t sm = 0;
    const int plans[4][2] = {{1, 0}, {1, 1}, {0, 0}, {0, 1}};           // {pipe, hits_in_global}
    for (int k = 0; k < 4; ++k) {
        if (plans[k][0] && (FILL || !h->render_pipe)) continue;
        c.pipe = plans[k][0]; c.hits_in_global = plans[k][1];
        sm = maze3d_smem_bytes(c, FILL);
        if (sm <= 227 * 1024) break;
    }
    if (sm > 227 * 1024) {
        mgb_set_error("3-D maze needs %zu bytes of shared memory per CTA (> 227 KB): reduce textures/resolution", sm);
        return MGB_ERR_ARG;
    }
    MazeArgs a2 = a;
    if (c.hits_in_global) {
        const size_t need = (size_t)grid * (c.pipe ? 2 : 1) * c.res_h * c.max_hits * sizeof(HitRec);
        if (need > h->hit_scratch_bytes) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
                mgb_set_error("maze renderer scratch must be allocated before stream capture: call reset() once first");
                return MGB_ERR_STATE;
            }
            MGB_CUDA(cudaStreamSynchronize(st));
            cudaFree(h->hit_scratch);
            h->hit_scratch = nullptr;
            MGB_CUDA(cudaMalloc(&h->hit_scratch, need));
            h->hit_scratch_bytes = need;
        }
        a2.hit_scratch = h->hit_scratch;
    }
    // The opt-in limit is a property of the kernel on a device, shared by every handle: each handle raises it once to the
    // device maximum (the same value from every handle and thread, so there is no ordering to get wrong and no global state).
    if (!h->render_attr_set[FILL ? 1 : 0]) {
        MGB_CUDA(maze_allow_max_dynamic_smem(maze3d_kernel<FILL>));
        h->render_attr_set[FILL ? 1 : 0] = 1;
    }
    h->smem3d = sm;
    maze3d_kernel<FILL><<<grid, kRenderThreads, sm, st>>>(c, a2);
    MGB_CUDA(cudaGetLastError());
    return MGB_OK;
}

// (Re)build the pose cache when tasks or textures changed: every free cell x 4 headings of every task is rendered once
// into its static layers.  Skipped (direct renderer used instead) when disabled or over the memory budget.
static int ensure_pose_cache(mgb_maze *h, cudaStream_t st)
{
    if (!h->cache_dirty) return MGB_OK;
    h->cache_ready = false;
    MazeConst &c = h->c;
    if (!h->cache_enabled || c.kind != MGB_MAZE_DISCRETE_3D || h->host_poses.empty()) { h->cache_dirty = false; return MGB_OK; }
    const size_t slots = h->host_poses.size(), px = (size_t)c.res_h * c.res_v;
    const double bytes = (double)slots * (px * (c.obs_dtype == MGB_OBS_U8 ? 8.0 : 12.0) + px / 4.0 + 16.0 +
                                          c.res_h * (1.0 + (double)c.max_hits * sizeof(HitRec)));
    if (bytes > h->cache_budget_gb * 1e9) { h->cache_dirty = false; h->cache_would_fit = false; return MGB_OK; }   // a decision, not a failure
    // From here on a failure (capture in progress, out of memory) leaves cache_dirty set: the next call retries instead of
    // silently rendering every frame with the slow direct renderer (round-1 advice).
    // cudaMalloc/cudaFree synchronise; a capture in progress cannot build the cache
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
        mgb_set_error("maze pose cache must be built before stream capture: call reset() once first");
        return MGB_ERR_STATE;
    }
    cudaFree(h->poses); cudaFree(h->pose_index); cudaFree(h->c_px); cudaFree(h->c_fid); cudaFree(h->c_colhits);
    cudaFree(h->c_hits); cudaFree(h->dyn); cudaFree(h->c_rgb8); cudaFree(h->c_gsig);
    cudaFree(h->c_px_all); cudaFree(h->c_fmask);
    cudaFree(h->c_vbase); cudaFree(h->c_var8); cudaFree(h->d_bake_desc); cudaFree(h->pose_rec);
    h->pose_rec = nullptr;
    h->c_vbase = nullptr; h->c_var8 = nullptr; h->d_bake_desc = nullptr; h->n_var_frames = 0;
    h->c_px_all = nullptr; h->c_fmask = nullptr;
    h->poses = nullptr; h->pose_index = nullptr; h->c_px = nullptr; h->c_fid = nullptr; h->c_colhits = nullptr;
    h->c_hits = nullptr; h->dyn = nullptr; h->c_rgb8 = nullptr; h->c_gsig = nullptr;
    MGB_CUDA(cudaMalloc(&h->poses, slots * sizeof(int4)));
    MGB_CUDA(cudaMalloc(&h->pose_index, h->host_pose_index.size() * sizeof(int32_t)));
    MGB_CUDA(cudaMalloc(&h->c_px, slots * px * sizeof(uint32_t)));
    MGB_CUDA(cudaMalloc(&h->c_fid, slots * px));
    MGB_CUDA(cudaMalloc(&h->c_colhits, slots * c.res_h));
    MGB_CUDA(cudaMalloc(&h->c_hits, slots * c.res_h * c.max_hits * sizeof(HitRec)));
    MGB_CUDA(cudaMalloc(&h->dyn, (size_t)h->n_pad * sizeof(EnvDyn)));
    MGB_CUDA(cudaMalloc(&h->c_rgb8, slots * px * 3));
    MGB_CUDA(cudaMalloc(&h->c_gsig, slots * ((px + 3) / 4)));
    MGB_CUDA(cudaMemsetAsync(h->c_gsig, 0xFF, slots * ((px + 3) / 4), st));
    MGB_CUDA(cudaMalloc(&h->c_fmask, slots * 2 * sizeof(uint64_t)));
    if (c.obs_dtype != MGB_OBS_U8) MGB_CUDA(cudaMalloc(&h->c_px_all, slots * px * sizeof(uint32_t)));
    MGB_CUDA(cudaMemcpy(h->poses, h->host_poses.data(), slots * sizeof(int4), cudaMemcpyHostToDevice));
    MGB_CUDA(cudaMemcpy(h->pose_index, h->host_pose_index.data(), h->host_pose_index.size() * sizeof(int32_t),
                        cudaMemcpyHostToDevice));
    MazeArgs a = maze_args(h);
    a.n = (int64_t)slots;
    a.do_step = 0;
    int rc = launch_render<true>(h, a, (unsigned)(slots < (size_t)h->num_sms ? slots : (size_t)h->num_sms), st);
    if (rc) return rc;
    // "all present" frames: which foods can change a pose's image at all, and the finished pixels with all of them
    // present (the compose kernel itself, fed one synthetic env per pose slot).  Screens the 4-pixel-group path cannot
    // take keep an all-ones mask, i.e. never use the baked frame.
    a.c_px_all = h->c_px_all; a.c_fmask = h->c_fmask;
    if ((c.res_v & 3) == 0 && (px & 127) == 0) {
        MGB_CUDA(cudaMemsetAsync(h->c_fmask, 0, slots * 2 * sizeof(uint64_t), st));
        maze3d_sig_kernel<<<(unsigned)slots, 256, 0, st>>>(c, a);
        MGB_CUDA(cudaGetLastError());
        if (h->c_px_all) MGB_CUDA(cudaMemcpyAsync(h->c_px_all, h->c_px, slots * px * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        // ---- variant frames: every pose whose image depends on k <= variant_bits foods gets its other 2^k - 1 finished
        // frames too (the all-visible one is c_rgb8[slot]); bits = the pose's foods in ascending slot order
        std::vector<BakeDesc> descs;
        if (h->variant_bits > 0 && c.obs_dtype == MGB_OBS_U8 && c.task_type == MGB_MAZE_SURVIVAL && (px * 3) % 16 == 0) {
            std::vector<uint64_t> fm(slots * 2);
            MGB_CUDA(cudaStreamSynchronize(st));
            MGB_CUDA(cudaMemcpy(fm.data(), h->c_fmask, slots * 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
            auto popc = [](uint64_t x) { int n = 0; while (x) { x &= x - 1; ++n; } return n; };
            for (int k = 0; k < 9; ++k) h->k_hist[k] = 0;
            for (size_t sl = 0; sl < slots; ++sl) {
                const int k = popc(fm[2 * sl]) + popc(fm[2 * sl + 1]);
                h->k_hist[k < 8 ? k : 8] += 1;
            }
            int bits = h->variant_bits;
            const double room = h->cache_budget_gb * 1e9 - bytes;
            size_t frames = 0;
            for (; bits > 0; --bits) {                     // largest k whose frames fit the cache budget
                frames = 0;
                for (size_t sl = 0; sl < slots; ++sl) {
                    const int k = popc(fm[2 * sl]) + popc(fm[2 * sl + 1]);
                    if (k >= 1 && k <= bits) frames += ((size_t)1 << k) - 1;
                }
                if ((double)frames * (double)(px * 3) <= room) break;
            }
            h->variant_bits_used = bits;
            if (bits > 0 && frames > 0) {
                std::vector<int32_t> vbase(slots, -1);
                descs.reserve(frames);
                for (size_t sl = 0; sl < slots; ++sl) {
                    const int k = popc(fm[2 * sl]) + popc(fm[2 * sl + 1]);
                    if (k < 1 || k > bits) continue;
                    vbase[sl] = (int32_t)descs.size();
                    for (int v = 0; v < (1 << k) - 1; ++v) {
                        BakeDesc bd;
                        bd.slot = (int32_t)sl; bd.pad = 0;
                        bd.present[0] = ~fm[2 * sl]; bd.present[1] = ~fm[2 * sl + 1];   // foods this pose never shows: irrelevant
                        int bit = 0;
                        for (int w = 0; w < 2; ++w) {
                            uint64_t m = fm[2 * sl + w];
                            while (m) {
                                const uint64_t low = m & (~m + 1);
                                m &= m - 1;
                                if ((v >> bit) & 1) bd.present[w] |= low;
                                ++bit;
                            }
                        }
                        descs.push_back(bd);
                    }
                }
                MGB_CUDA(cudaMalloc(&h->c_vbase, slots * sizeof(int32_t)));
                MGB_CUDA(cudaMalloc(&h->c_var8, descs.size() * px * 3));
                MGB_CUDA(cudaMalloc(&h->d_bake_desc, descs.size() * sizeof(BakeDesc)));
                MGB_CUDA(cudaMemcpy(h->c_vbase, vbase.data(), slots * sizeof(int32_t), cudaMemcpyHostToDevice));
                MGB_CUDA(cudaMemcpy(h->d_bake_desc, descs.data(), descs.size() * sizeof(BakeDesc), cudaMemcpyHostToDevice));
                h->n_var_frames = (int64_t)descs.size();
                MazeArgs av = a;
                av.c_var8 = h->c_var8; av.bake_desc = h->d_bake_desc;
                maze3d_varinit_kernel<<<(unsigned)descs.size(), 256, 0, st>>>(c, av);      // static colours first ...
                MGB_CUDA(cudaGetLastError());
            }
        }
        a.bake = 1;
        a.do_parts = 1;
        maze3d_compose_kernel<<<(unsigned)slots, kComposeThreads, 0, st>>>(c, a);
        MGB_CUDA(cudaGetLastError());
        if (!descs.empty()) {                               // ... then each variant's tints (compose kernel, bake mode)
            MazeArgs av = a;
            av.n = (int64_t)descs.size();
            av.c_rgb8 = h->c_var8; av.bake_desc = h->d_bake_desc;
            maze3d_compose_kernel<<<(unsigned)descs.size(), kComposeThreads, 0, st>>>(c, av);
            MGB_CUDA(cudaGetLastError());
            h->launches += 2;
        }
        a.bake = 0;
        h->launches += 2;
    } else {
        MGB_CUDA(cudaMemsetAsync(h->c_fmask, 0xFF, slots * 2 * sizeof(uint64_t), st));
    }
    {   // merged per-pose records for the step logic (after c_fmask and c_vbase are final)
        const int64_t count = (int64_t)h->host_pose_index.size();
        MGB_CUDA(cudaMalloc(&h->pose_rec, (size_t)count * sizeof(PoseRec)));
        maze_pose_rec_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(h->pose_index, h->c_fmask, h->c_vbase,
                                                                            reinterpret_cast<PoseRec *>(h->pose_rec), count);
        MGB_CUDA(cudaGetLastError());
    }
    MGB_CUDA(cudaStreamSynchronize(st));
    h->n_poses = (int64_t)slots;
    h->cache_bytes = bytes + (double)h->n_var_frames * (double)(px * 3);
    h->cache_ready = true;
    h->cache_dirty = false;
    h->launches += 1;
    return MGB_OK;
}

static int launch_observe(mgb_maze *h, MazeArgs &a, cudaStream_t st)
{
    const MazeConst &c = h->c;
    if (c.kind == MGB_MAZE_2D) {
        const int W = 2 * c.view_grid + 1;
        const size_t sm = (size_t)k2dThreads * W * W * 4;
        if (sm > 48 * 1024 && sm > h->m2d_smem_set) {      // view_grid >= 5: above the default dynamic shared-memory limit
            MGB_REQUIRE(sm <= 200 * 1024, "view_grid too large for the 2-D observation tile");
            MGB_CUDA(maze_allow_max_dynamic_smem(maze2d_kernel));   // the device maximum: the same value from every handle
            h->m2d_smem_set = sm;
        }
        maze2d_kernel<<<(unsigned)((h->n + k2dThreads - 1) / k2dThreads), k2dThreads, sm, st>>>(c, a);
    } else {
        int rc = ensure_pose_cache(h, st);
        if (rc) return rc;
        if (h->cache_ready) {
            // memoised path: integer step logic, then compose static pose layers with the current food state
            a.poses = h->poses; a.pose_index = h->pose_index; a.c_px = h->c_px; a.c_fid = h->c_fid;
            a.c_colhits = h->c_colhits; a.c_hits = h->c_hits; a.dyn = h->dyn; a.c_rgb8 = h->c_rgb8; a.c_gsig = h->c_gsig;
            a.c_px_all = h->c_px_all; a.c_fmask = h->c_fmask;
            a.c_vbase = h->c_vbase; a.c_var8 = h->c_var8; a.pose_rec = h->pose_rec;   // (re)built by ensure_pose_cache just above
            // uint8 frames whose columns are whole 16-pixel runs: ONE fused launch (logic + TMA-moved frame)
            const size_t frame_bytes = (size_t)c.res_h * c.res_v * 3;
            if (h->fused_step && c.obs_dtype == MGB_OBS_U8 && (c.res_v & 15) == 0 && ((size_t)c.res_h * c.res_v) % 128 == 0 &&
                frame_bytes <= (size_t)64 * kStepChunkPx * 3 &&
                (reinterpret_cast<uintptr_t>(a.obs) & 15u) == 0) {
                const size_t ring_bytes = (size_t)kStepSlots * kStepChunkPx * 3;       // 96 KB: two CTAs per SM
                if (h->step_smem_set == 0) {
                    MGB_CUDA(maze_allow_max_dynamic_smem(maze3d_step_kernel));
                    h->step_smem_set = ring_bytes;
                }
                const int64_t grid = (int64_t)h->num_sms * 2;
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3((unsigned)(h->n < grid ? h->n : grid));
                cfg.blockDim = dim3(kStepThreads);
                cfg.dynamicSmemBytes = ring_bytes;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = h->step_pdl ? 1 : 0;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                MGB_CUDA(cudaLaunchKernelEx(&cfg, maze3d_step_kernel, c, a));
                MGB_CUDA(cudaGetLastError());
                h->launches += 1;
                return MGB_OK;
            }
            maze3d_logic_kernel<true><<<(unsigned)((h->n + 127) / 128), 128, 0, st>>>(c, a);
            MGB_CUDA(cudaGetLastError());
            {
                int &ctas_per_sm = h->compose_ctas_per_sm;
                if (!ctas_per_sm) {
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, maze3d_compose_kernel, kComposeThreads, 0) !=
                            cudaSuccess || ctas_per_sm < 1) ctas_per_sm = 4;
                }
                const int64_t resident = (int64_t)h->num_sms * ctas_per_sm;
                int64_t parts = (4 * resident + h->n - 1) / h->n;  // aim at >= 4 work items per resident CTA
                parts = parts < 1 ? 1 : (parts > 4 ? 4 : parts);
                a.do_parts = (int)parts;
                int64_t items = h->n * parts;
                int &persistent = h->compose_persistent;
                if (persistent < 0) { const char *ev = getenv("MGB_COMPOSE_PERSISTENT"); persistent = ev ? atoi(ev) : 0; }   // measured: plain grid 65/295 us vs persistent 72/319 us (1024/8192 envs)
                if (!persistent) items = items < resident ? items : 0x7fffffff;   // plain grid: one CTA per item
                if (!persistent) { maze3d_compose_kernel<<<(unsigned)(h->n * parts), kComposeThreads, 0, st>>>(c, a); }
                else
                maze3d_compose_kernel<<<(unsigned)(items < resident ? items : resident), kComposeThreads, 0, st>>>(c, a);
            }
            h->launches += 1;
        } else {
            if (a.do_step && c.kind == MGB_MAZE_DISCRETE_3D) {     // logic for all envs in parallel, then render only
                maze3d_logic_kernel<false><<<(unsigned)((h->n + 127) / 128), 128, 0, st>>>(c, a);
                MGB_CUDA(cudaGetLastError());
                h->launches += 1;
                a.do_step = 0;
            }
            rc = launch_render<false>(h, a, (unsigned)(h->n < h->num_sms ? h->n : h->num_sms), st);
            if (rc) return rc;
        }
    }
    MGB_CUDA(cudaGetLastError());
    h->launches += 1;
    return MGB_OK;
}

extern "C" int mgb_maze_reset(mgb_maze *h, const uint8_t *mask_dev, void *obs_dev, void *stream)
{
    MgbRange nvtx_range("mgb_maze_reset");
    MGB_REQUIRE(h, "null handle");
    int rc = maze_ready(h);
    if (rc) return rc;
    MgbDeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    MazeArgs a = maze_args(h);
    a.mask = mask_dev;
    maze_reset_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, st>>>(h->c, a);
    MGB_CUDA(cudaGetLastError());
    h->launches += 1;
    if (obs_dev) {
        a.obs = obs_dev; a.do_step = 0; a.mask = nullptr;
        return launch_observe(h, a, st);
    }
    return MGB_OK;
}

extern "C" int mgb_maze_rollout(mgb_maze *h, int32_t T, const int32_t *act_dev, uint64_t act_seed, int32_t *act_out_dev,
                                void *obs_dev, double *rew_dev, uint8_t *done_dev, void *stream)
{
    MgbRange nvtx_range("mgb_maze_rollout");
    MGB_REQUIRE(h, "null handle");
    MGB_REQUIRE(T > 0, "T must be positive");
    MGB_REQUIRE(h->c.kind == MGB_MAZE_2D || h->c.kind == MGB_MAZE_DISCRETE_3D,
                "mgb_maze_rollout serves MetaMaze2D and MetaMazeDiscrete3D");
    int rc = maze_ready(h);
    if (rc) return rc;
    MgbDeviceGuard guard(h->device);
    MazeArgs a = maze_args(h);
    a.act = act_dev; a.obs = obs_dev; a.rew = rew_dev; a.done = done_dev; a.do_step = 1;
    a.T = T; a.act_seed = act_seed; a.t_base = h->t_base; a.act_out = act_out_dev;
    a.mir = h->mir;
    if (h->mir.count != 0)
        MGB_REQUIRE(h->mir_win.holds(obs_dev, (uint64_t)T * h->n * (uint64_t)mgb_maze_obs_bytes_per_env(h)) &&
                        h->mir_win.holds(rew_dev, (uint64_t)T * h->n * 8) && h->mir_win.holds(done_dev, (uint64_t)T * h->n) &&
                        h->mir_win.holds(act_out_dev, (uint64_t)T * h->n * 4),
                    "mirrors are on but an output lies outside the mirrored arena (set_mirrors([]) first)");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->c.kind == MGB_MAZE_DISCRETE_3D) {
        MGB_REQUIRE(h->mir.count == 0, "output mirrors are implemented for the MetaMaze2D rollout only");
        rc = ensure_pose_cache(h, st);
        if (rc) return rc;
        MGB_REQUIRE(h->cache_ready, "the fused 3-D rollout runs on the pose cache (MGB_MAZE_CACHE=0 or cache budget too small)");
        a.poses = h->poses; a.pose_index = h->pose_index; a.c_px = h->c_px; a.c_fid = h->c_fid;
        a.c_colhits = h->c_colhits; a.c_hits = h->c_hits; a.dyn = h->dyn; a.c_rgb8 = h->c_rgb8; a.c_gsig = h->c_gsig;
        a.c_px_all = h->c_px_all; a.c_fmask = h->c_fmask;
        a.c_vbase = h->c_vbase; a.c_var8 = h->c_var8; a.pose_rec = h->pose_rec;
        a.bake = 0;
        const int64_t resident = (int64_t)h->num_sms * 5;          // __launch_bounds__(256, 5): 48 registers, no spills
        const size_t qbytes = ((size_t)h->c.res_h * h->c.res_v / 4 + 1) * sizeof(int);
        MGB_REQUIRE(qbytes <= 200 * 1024, "screen too large for the fused rollout's group queue");
        if (qbytes > 40 * 1024)
            MGB_CUDA(maze_allow_max_dynamic_smem(maze3d_rollout_kernel));
        maze3d_rollout_kernel<<<(unsigned)(h->n < resident ? h->n : resident), kComposeThreads, qbytes, st>>>(h->c, a);
        MGB_CUDA(cudaGetLastError());
        h->t_base += (uint32_t)T;
        h->launches += 1;
        return MGB_OK;
    }
    const int W = 2 * h->c.view_grid + 1;
    const size_t sm = (size_t)2 * k2dThreads * W * W * 4;
    const unsigned blocks = (unsigned)((h->n + k2dThreads - 1) / k2dThreads);
    if (sm > 48 * 1024) {
        MGB_CUDA(maze_allow_max_dynamic_smem(maze2d_rollout_kernel<0>));
        MGB_CUDA(maze_allow_max_dynamic_smem(maze2d_rollout_kernel<1>));
        MGB_CUDA(maze_allow_max_dynamic_smem(maze2d_rollout_kernel<2>));
    }
    if (h->mir.count == MGB_MIRROR_MULTICAST) {
        MGB_REQUIRE(h->n % 4 == 0, "multicast outputs need num_envs % 4 == 0");
        MGB_REQUIRE((((uintptr_t)done_dev | (uintptr_t)obs_dev | (uintptr_t)act_out_dev) & 3) == 0 && ((uintptr_t)rew_dev & 7) == 0,
                    "multicast outputs must be 4-byte (rewards: 8-byte) aligned");
        maze2d_rollout_kernel<2><<<blocks, k2dThreads, sm, st>>>(h->c, a);
    } else if (h->mir.count > 0) maze2d_rollout_kernel<1><<<blocks, k2dThreads, sm, st>>>(h->c, a);
    else maze2d_rollout_kernel<0><<<blocks, k2dThreads, sm, st>>>(h->c, a);
    MGB_CUDA(cudaGetLastError());
    h->t_base += (uint32_t)T;
    h->launches += 1;
    return MGB_OK;
}

extern "C" int mgb_maze_set_mirrors(mgb_maze *h, int count, const int64_t *byte_delta)
{
    MGB_REQUIRE(h, "null handle");
    MGB_REQUIRE(count >= 0 && count <= MGB_MAX_MIRRORS && (count == 0 || byte_delta), "count out of range");
    MgbMirrors m = {};
    for (int i = 0; i < count; ++i) {
        MGB_REQUIRE((byte_delta[i] & 15) == 0, "mirror deltas must be multiples of 16 bytes");
        m.delta[i] = byte_delta[i];
    }
    m.count = count;
    h->mir = m;
    return MGB_OK;
}

extern "C" int mgb_maze_set_mirror_window(mgb_maze *h, const void *base, uint64_t bytes)
{
    MGB_REQUIRE(h, "null handle");
    h->mir_win.base = reinterpret_cast<uintptr_t>(base);
    h->mir_win.bytes = bytes;
    return MGB_OK;
}

extern "C" int mgb_maze_set_multicast(mgb_maze *h, int64_t byte_delta)
{
    MGB_REQUIRE(h, "null handle");
    MGB_REQUIRE((byte_delta & 15) == 0, "multicast delta must be a multiple of 16 bytes");
    MgbMirrors m = {};
    if (byte_delta != 0) { m.count = MGB_MIRROR_MULTICAST; m.delta[0] = byte_delta; }
    h->mir = m;
    return MGB_OK;
}

extern "C" int mgb_maze_step_continuous(mgb_maze *h, const float *act_dev, void *obs_dev, double *rew_dev,
                                        uint8_t *done_dev, void *stream)
{
    MgbRange nvtx_range("mgb_maze_step_continuous");
    MGB_REQUIRE(h && act_dev && obs_dev && rew_dev && done_dev, "null argument");
    MGB_REQUIRE(h->c.kind == MGB_MAZE_CONTINUOUS_3D, "mgb_maze_step_continuous needs a MGB_MAZE_CONTINUOUS_3D handle");
    int rc = maze_ready(h);
    if (rc) return rc;
    MgbDeviceGuard guard(h->device);
    MazeArgs a = maze_args(h);
    a.act_c = act_dev; a.obs = obs_dev; a.rew = rew_dev; a.done = done_dev; a.do_step = 1;
    return launch_observe(h, a, (cudaStream_t)stream);
}

__global__ void maze_pose_kernel(MazeArgs a, float *pos_out, double *ori_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    const float2 p = a.cpos[e];
    pos_out[2 * e] = p.x; pos_out[2 * e + 1] = p.y;
    if (ori_out) ori_out[e] = a.cori[e];
}

extern "C" int mgb_maze_pose(mgb_maze *h, float *pos_dev, double *ori_dev, void *stream)
{
    MGB_REQUIRE(h && pos_dev, "null argument");
    MGB_REQUIRE(h->c.kind == MGB_MAZE_CONTINUOUS_3D, "mgb_maze_pose needs a MGB_MAZE_CONTINUOUS_3D handle");
    MgbDeviceGuard guard(h->device);
    MazeArgs a = maze_args(h);
    maze_pose_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, pos_dev, ori_dev);
    MGB_CUDA(cudaGetLastError());
    h->launches += 1;
    return MGB_OK;
}

extern "C" int mgb_maze_step(mgb_maze *h, const int32_t *act_dev, void *obs_dev, double *rew_dev, uint8_t *done_dev,
                             void *stream)
{
    MgbRange nvtx_range("mgb_maze_step");
    MGB_REQUIRE(h && act_dev && obs_dev && rew_dev && done_dev, "null argument");
    MGB_REQUIRE(h->c.kind != MGB_MAZE_CONTINUOUS_3D, "use mgb_maze_step_continuous for the continuous maze");
    int rc = maze_ready(h);
    if (rc) return rc;
    MgbDeviceGuard guard(h->device);
    MazeArgs a = maze_args(h);
    a.act = act_dev; a.obs = obs_dev; a.rew = rew_dev; a.done = done_dev; a.do_step = 1;
    return launch_observe(h, a, (cudaStream_t)stream);
}

extern "C" int mgb_maze_state(mgb_maze *h, int32_t *agent_dev, double *life_dev, void *stream)
{
    MGB_REQUIRE(h && agent_dev, "null argument");
    MgbDeviceGuard guard(h->device);
    MazeArgs a = maze_args(h);
    maze_state_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, agent_dev, life_dev);
    MGB_CUDA(cudaGetLastError());
    h->launches += 1;
    return MGB_OK;
}
