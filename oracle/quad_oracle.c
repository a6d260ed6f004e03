/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Three precisions of the quadrotor oracle, see quad_oracle_impl.h.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: numpy does not fuse multiply-add in its scalar/ufunc loops).
 */
#include <math.h>
#include "quad_oracle.h"

#define QO_CAT2(a, b) a##b
#define QO_CAT(a, b) QO_CAT2(a, b)

/* ---- all float32: a simulator that was never reset() (velocity-task generator) ---- */
#define T float
#define TV float
#define QO_NAME(x) QO_CAT(qo_f32_, x)
#define QO_SQRTT sqrtf
#define QO_SQRTV sqrtf
#define QO_ATAN2T atan2f
#include "quad_oracle_impl.h"
#undef T
#undef TV
#undef QO_NAME
#undef QO_SQRTT
#undef QO_SQRTV
#undef QO_ATAN2T

/* ---- mixed: float32 state with float64 velocity vectors (after reset()) ---- */
#define T float
#define TV double
#define QO_NAME(x) QO_CAT(qo_mix_, x)
#define QO_SQRTT sqrtf
#define QO_SQRTV sqrt
#define QO_ATAN2T atan2f
#include "quad_oracle_impl.h"
#undef T
#undef TV
#undef QO_NAME
#undef QO_SQRTT
#undef QO_SQRTV
#undef QO_ATAN2T

/* ---- all float64: arbiter ---- */
#define T double
#define TV double
#define QO_NAME(x) QO_CAT(qo_f64_, x)
#define QO_SQRTT sqrt
#define QO_SQRTV sqrt
#define QO_ATAN2T atan2
#include "quad_oracle_impl.h"
