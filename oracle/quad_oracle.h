/* TEST INFRASTRUCTURE -- interface of the CPU oracle for the quadrotor path (see quad_oracle_impl.h). */
#ifndef QUAD_ORACLE_H
#define QUAD_ORACLE_H
typedef struct {
    double h;        /* "precision", config.json:2  */
    double m;        /* "quality" (mass), config.json:3 */
    double Iinv[9];  /* np.linalg.inv(float32 inertia) as computed by the host (quadrotorsim.py:64) */
    double Dm[3], Df[3];
    double cg[3];
    double ct0, ct1, ct2, mm, jm, phi, ra;
    double fail_v, fail_r, fail_w;
    double prop[12];
    double lm[4];    /* np.linalg.norm(float32 propeller coord) (quadrotorsim.py:146) */
    double vmin, vmax;
} qo_cfg;
#endif
