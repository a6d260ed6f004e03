// placeholder until the maze kernels land (this file is replaced in the next milestone)
#include "mgb_common.cuh"
#define NYI() do { mgb_set_error("%s: maze path not built yet", __func__); return MGB_ERR_STATE; } while (0)
extern "C" int mgb_maze_create(mgb_maze **, int64_t, const mgb_maze_cfg *, int, int64_t) { NYI(); }
extern "C" void mgb_maze_destroy(mgb_maze *) {}
extern "C" int64_t mgb_maze_obs_bytes_per_env(const mgb_maze *) { return -1; }
extern "C" int mgb_maze_set_textures(mgb_maze *, const uint8_t *, int32_t, const uint8_t *, int32_t) { NYI(); }
extern "C" int mgb_maze_set_task(mgb_maze *, int32_t, const int8_t *, const int8_t *, const double *, const int32_t *,
                                 const mgb_maze_task_scalars *, const int32_t *) { NYI(); }
extern "C" int mgb_maze_reset(mgb_maze *, const uint8_t *, void *, void *) { NYI(); }
extern "C" int mgb_maze_step(mgb_maze *, const int32_t *, void *, double *, uint8_t *, void *) { NYI(); }
extern "C" int mgb_maze_set_options(mgb_maze *, int) { NYI(); }
extern "C" int mgb_maze_state(mgb_maze *, int32_t *, double *, void *) { NYI(); }
extern "C" int64_t mgb_maze_launch_count(const mgb_maze *) { return -1; }
