/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement of the reference quadrotor hot path, written from the equations of
 *   metagym/quadrotor/quadrotorsim.py:122-221 (_run_internal, _check_failure)
 *   metagym/quadrotor/quadrotorsim.py:260-304 (get_state, get_sensor, step)
 *   metagym/quadrotor/env.py:127-165,211-281   (Quadrotor.step, _get_reward, _check_collision, _update_state)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * This header is included three times by quad_oracle.c with different scalar types:
 *   T   = storage/compute type of position, rotation matrix, propeller speed and all "scalar" rotor math
 *   TV  = storage type of global velocity and body angular velocity
 * The reference is mixed precision by accident: after reset() the velocity vectors are
 * float64 numpy arrays while everything else is float32, and python-float constants are "weak" (numpy >= 2, NEP 50),
 * i.e. they take the type of the numpy operand.  C's usual arithmetic conversions reproduce numpy's array
 * promotion (float (x) double -> double); the weak constants are reproduced by the explicit KT()/KV() casts.
 *   _f32 : T=float,  TV=float   (state never reset(): e.g. the velocity-task generator, quadrotorsim.py:306)
 *   _mix : T=float,  TV=double  (state after reset(): what Quadrotor.step sees in normal use)
 *   _f64 : T=double, TV=double  (arbiter; not a reference mode)
 */

#define KT(c) ((T)(c))   /* python float meeting a T-typed numpy value  */
#define KV(c) ((TV)(c))  /* python float meeting a TV-typed numpy value */

static void QO_NAME(substep)(const qo_cfg *c, const T vclamp[4], T p[3], TV v[3], TV om[3], T w[4], T R[9], T Ri[9],
                             T *power_out)
{
    /* body velocity, shared by the four rotors (quadrotorsim.py:147-148) */
    TV bv[3];
    for (int r = 0; r < 3; ++r) bv[r] = Ri[3 * r] * v[0] + Ri[3 * r + 1] * v[1] + Ri[3 * r + 2] * v[2];

    T me[4], pw[4];
    T fz = KT(0.0);
    T tq[3] = {KT(0.0), KT(0.0), KT(0.0)};
    for (int i = 0; i < 4; ++i) {
        const T px = KT(c->prop[3 * i]), py = KT(c->prop[3 * i + 1]), pz = KT(c->prop[3 * i + 2]);
        const T V = vclamp[i];
        T phi_w = KT(c->phi) * w[i];                                 /* :136 */
        me[i] = KT(c->phi / c->ra) * (V - phi_w);                    /* :137-138 */
        T q = me[i] / KT(c->phi) * V;                                /* :139 */
        pw[i] = q < 0 ? -q : q;
        T dw = KT(1.0 / c->jm) * (me[i] - KT(c->mm));                /* :141-142 */
        T wm = w[i] + KT(c->h) * dw;                                 /* :144-145 */
        /* (omega x p_i) * |p_i| , z component only is used           :149-151 */
        TV cz = om[0] * py - om[1] * px;
        TV v1 = bv[2] + cz * KT(c->lm[i]);
        (void)pz;
        TV sgn = v1 > 0 ? KV(1.0) : KV(-1.0);
        /* :154-156 ; first term stays in T, the rest is promoted by v1 */
        TV thrust = (KT(c->ct0) * wm * wm) + (KT(c->ct1) * wm) * v1 + (KV(c->ct2) * v1) * v1 * sgn;
        w[i] = wm;
        fz = (T)(fz + thrust);                                       /* :159 */
        T th = (T)thrust;                                            /* :160-162: -(0,0,th) x p_i */
        T nx = KT(-0.0), ny = KT(-0.0), nz = -th;
        tq[0] += ny * pz - nz * py;
        tq[1] += nz * px - nx * pz;
        tq[2] += nx * py - ny * px;
    }
    tq[2] += -me[0] + me[1] - me[2] + me[3];                          /* :164 */

    /* drag :166-172 */
    TV vn = (TV)QO_SQRTV(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    TV on = (TV)QO_SQRTV(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    TV fdrag[3], tdrag[3];
    for (int r = 0; r < 3; ++r) {
        T d = KT(c->Df[r]);
        TV s = (d * Ri[3 * r]) * v[0] + (d * Ri[3 * r + 1]) * v[1] + (d * Ri[3 * r + 2]) * v[2];
        fdrag[r] = -vn * s;
        tdrag[r] = -on * (KT(c->Dm[r]) * om[r]);
    }
    /* gravity :174-178 */
    T fg[3], tg[3];
    for (int r = 0; r < 3; ++r) fg[r] = (Ri[3 * r + 2] * KT(-9.8)) * KT(c->m);
    {
        const T gx = KT(c->cg[0]), gy = KT(c->cg[1]), gz = KT(c->cg[2]);
        tg[0] = -(fg[1] * gz - fg[2] * gy);
        tg[1] = -(fg[2] * gx - fg[0] * gz);
        tg[2] = -(fg[0] * gy - fg[1] * gx);
    }
    TV F[3], Tq[3];
    F[0] = (KT(0.0) + fg[0]) + fdrag[0];
    F[1] = (KT(0.0) + fg[1]) + fdrag[1];
    F[2] = (fz + fg[2]) + fdrag[2];
    for (int r = 0; r < 3; ++r) Tq[r] = (tq[r] + tg[r]) + tdrag[r];

    /* translation :183-187 */
    TV ba[3], a[3];
    for (int r = 0; r < 3; ++r) ba[r] = F[r] / KV(c->m);
    for (int r = 0; r < 3; ++r) a[r] = R[3 * r] * ba[0] + R[3 * r + 1] * ba[1] + R[3 * r + 2] * ba[2];
    for (int r = 0; r < 3; ++r) {
        p[r] = (T)(p[r] + (v[r] * KV(c->h) + KV(0.5 * c->h * c->h) * a[r]));
        v[r] = v[r] + KV(c->h) * a[r];
    }
    *power_out = ((pw[0] + pw[1]) + pw[2]) + pw[3];                   /* :188 */

    /* rotation :190-204 */
    TV al[3];
    for (int r = 0; r < 3; ++r)
        al[r] = KT(c->Iinv[3 * r]) * Tq[0] + KT(c->Iinv[3 * r + 1]) * Tq[1] + KT(c->Iinv[3 * r + 2]) * Tq[2];
    T wt[3];
    for (int r = 0; r < 3; ++r) wt[r] = (T)(om[r] + KV(0.5 * c->h) * al[r]);
    T RS[9];
    for (int r = 0; r < 3; ++r) {
        const T r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
        RS[3 * r + 0] = r1 * wt[2] + r2 * (-wt[1]);
        RS[3 * r + 1] = r0 * (-wt[2]) + r2 * wt[0];
        RS[3 * r + 2] = r0 * wt[1] + r1 * (-wt[0]);
    }
    for (int k = 0; k < 9; ++k) R[k] = R[k] + KT(c->h) * RS[k];
    for (int r = 0; r < 3; ++r) om[r] = om[r] + KV(c->h) * al[r];

    /* general 3x3 inverse (the reference calls LAPACK; R is NOT orthogonal, so this is a real inverse) :206-208 */
    {
        const T a00 = R[0], a01 = R[1], a02 = R[2], a10 = R[3], a11 = R[4], a12 = R[5], a20 = R[6], a21 = R[7],
                a22 = R[8];
        const T c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
        const T det = a00 * c00 + a01 * c01 + a02 * c02;
        const T id = KT(1.0) / det;
        Ri[0] = c00 * id;
        Ri[1] = (a02 * a21 - a01 * a22) * id;
        Ri[2] = (a01 * a12 - a02 * a11) * id;
        Ri[3] = c01 * id;
        Ri[4] = (a00 * a22 - a02 * a20) * id;
        Ri[5] = (a02 * a10 - a00 * a12) * id;
        Ri[6] = c02 * id;
        Ri[7] = (a01 * a20 - a00 * a21) * id;
        Ri[8] = (a00 * a11 - a01 * a10) * id;
    }
}

static void QO_NAME(invert)(const T R[9], T Ri[9])
{
    const T a00 = R[0], a01 = R[1], a02 = R[2], a10 = R[3], a11 = R[4], a12 = R[5], a20 = R[6], a21 = R[7],
            a22 = R[8];
    const T c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const T det = a00 * c00 + a01 * c01 + a02 * c02;
    const T id = KT(1.0) / det;
    Ri[0] = c00 * id;
    Ri[1] = (a02 * a21 - a01 * a22) * id;
    Ri[2] = (a01 * a12 - a02 * a11) * id;
    Ri[3] = c01 * id;
    Ri[4] = (a00 * a22 - a02 * a20) * id;
    Ri[5] = (a02 * a10 - a00 * a12) * id;
    Ri[6] = c02 * id;
    Ri[7] = (a01 * a20 - a00 * a21) * id;
    Ri[8] = (a00 * a11 - a01 * a10) * id;
}

/* _check_failure, quadrotorsim.py:212-221.  0 ok, 1 range, 2 velocity, 3 angular velocity */
static int QO_NAME(fail)(const qo_cfg *c, const T p[3], const TV v[3], const TV om[3])
{
    if ((T)QO_SQRTT(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) > KT(c->fail_r)) return 1;
    if ((TV)QO_SQRTV(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) > KV(c->fail_v)) return 2;
    if ((TV)QO_SQRTV(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]) > KV(c->fail_w)) return 3;
    return 0;
}

/*
 * One env.step() for n independent envs (env.py:127-165).
 *   state  [n][22]  p3 v3 om3 w4 R9 (row-major R), doubles on the interface (exact carriers of the T/TV values)
 *   ct     [n]      episode step counter (env.py:65,128,150,161)
 *   act    [n][4]   float32 actions
 *   tgt    [n_tasks][nt][3] velocity targets (task 2 only), env2task[n]
 *   obs    [n][obs_dim] float32, rew [n] double, done [n], fail [n], power [n]
 * task: 0 no_collision, 1 hovering_control, 2 velocity_control
 */
void QO_NAME(env_step)(const qo_cfg *c, int n, double *state, int *ct, const float *act, int task, double dt,
                       int nt, double healthy, const float *tgt, const int *env2task, float *obs, double *rew,
                       unsigned char *done, int *fail, double *power)
{
    const int substeps = (int)(dt / c->h);                           /* quadrotorsim.py:302 */
    const int obs_dim = task == 2 ? 19 : 16;
    for (int e = 0; e < n; ++e) {
        double *s = state + 22 * (long)e;
        T p[3], w[4], R[9], Ri[9];
        TV v[3], om[3];
        for (int k = 0; k < 3; ++k) { p[k] = (T)s[k]; v[k] = (TV)s[3 + k]; om[k] = (TV)s[6 + k]; }
        for (int k = 0; k < 4; ++k) w[k] = (T)s[9 + k];
        for (int k = 0; k < 9; ++k) R[k] = (T)s[13 + k];
        QO_NAME(invert)(R, Ri);
        ct[e] += 1;                                                  /* env.py:128 */
        const double zoff = task == 2 ? 0.0 : 5.0;                   /* env.py:97,112: z_offset only with a map */
        const T z_old = p[2] + KT(zoff);                             /* env.py:131-133 */
        const double x_old = (double)p[0] + qo_xoff, y_old = (double)p[1] + qo_yoff;
        T vclamp[4];
        for (int k = 0; k < 4; ++k) {                                /* quadrotorsim.py:130-134 */
            double a = (double)act[4 * (long)e + k];
            if (a > c->vmax) a = c->vmax; else if (a < c->vmin) a = c->vmin;
            vclamp[k] = KT(a);
        }
        T pwr = KT(0.0);
        int fc = 0;
        for (int k = 0; k < substeps; ++k) {
            QO_NAME(substep)(c, vclamp, p, v, om, w, R, Ri, &pwr);
            fc = QO_NAME(fail)(c, p, v, om);
            if (fc) break;   /* the reference raises here; the batched engine freezes the env and reports done */
        }
        for (int k = 0; k < 3; ++k) { s[k] = (double)p[k]; s[3 + k] = (double)v[k]; s[6 + k] = (double)om[k]; }
        for (int k = 0; k < 4; ++k) s[9 + k] = (double)w[k];
        for (int k = 0; k < 9; ++k) s[13 + k] = (double)R[k];

        /* get_state / get_sensor, quadrotorsim.py:260-293 */
        TV bvel[3];
        T bpos[3], acc[3];
        for (int r = 0; r < 3; ++r) {
            bvel[r] = Ri[3 * r] * v[0] + Ri[3 * r + 1] * v[1] + Ri[3 * r + 2] * v[2];
            bpos[r] = Ri[3 * r] * p[0] + Ri[3 * r + 1] * p[1] + Ri[3 * r + 2] * p[2];
            acc[r] = KT(0.0) + Ri[3 * r + 2] * KT(-9.8);
        }
        const T roll = (T)QO_ATAN2T(R[7], R[8]);
        const T pitch = (T)QO_ATAN2T(-R[6], (T)QO_SQRTT(R[7] * R[7] + R[8] * R[8]));
        const T yaw = (T)QO_ATAN2T(R[3], R[0]);
        float *o = obs + (long)obs_dim * e;                          /* env.py:193-209 key order */
        o[0] = (float)bvel[0]; o[1] = (float)bvel[1]; o[2] = (float)bvel[2];
        o[3] = (float)bpos[0]; o[4] = (float)bpos[1]; o[5] = (float)bpos[2];
        o[6] = (float)acc[0];  o[7] = (float)acc[1];  o[8] = (float)acc[2];
        o[9] = (float)om[0];   o[10] = (float)om[1];  o[11] = (float)om[2];
        o[12] = (float)pitch;  o[13] = (float)roll;   o[14] = (float)yaw;
        o[15] = (float)(p[2] + KT(zoff));
        const float *trow = 0;
        if (task == 2) {
            trow = tgt + ((long)env2task[e] * nt) * 3;
            int t = ct[e] < nt - 1 ? ct[e] : nt - 1;                 /* env.py:270-274 */
            o[16] = trow[3 * t]; o[17] = trow[3 * t + 1]; o[18] = trow[3 * t + 2];
        }

        /* reward / done, env.py:144-161,211-260 */
        const T z_new = p[2] + KT(zoff);
        const T zmin = z_old < z_new ? z_old : z_new;
        int collision = (task != 2) && (zmin < KT(0.0));             /* flat map: z < np.any(zeros) == z < False */
        if (task != 2 && qo_sat) {                                   /* env.py:248-260 with an obstacle map */
            const T zmax = z_old > z_new ? z_old : z_new;
            const int any = qo_any_obstacle(x_old, y_old, (double)p[0] + qo_xoff, (double)p[1] + qo_yoff);
            collision = ((int)QO_FLOORT(zmin) < any) || ((int)QO_CEILT(zmax) < any);
        }
        T e_cost = KT(dt) * pwr;
        double r = -(double)(e_cost < KT(healthy) ? e_cost : KT(healthy));
        int dn = 0;
        if (task == 0) {
            r += collision ? 0.0 : healthy;
        } else if (task == 1) {
            double tr = collision ? 0.0 : healthy;
            TV vn = (TV)QO_SQRTV(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            TV on = (TV)QO_SQRTV(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
            tr -= 1.0 * vn + 1.0 * on;
            T zm = KT(0.0) - p[2];
            if (zm < 0) zm = -zm;
            if (zm < KT(0.5)) tr += 10; else { T q = KT(0.5) - zm; tr += (q > KT(-20) ? q : KT(-20)); }
            r += tr;
        } else {
            const float *g = trow + 3 * (ct[e] - 1);                 /* env.py:153 */
            T bt[3];
            for (int q = 0; q < 3; ++q) bt[q] = Ri[3 * q] * (T)g[0] + Ri[3 * q + 1] * (T)g[1] + Ri[3 * q + 2] * (T)g[2];
            TV d0 = bt[0] - bvel[0], d1 = bt[1] - bvel[1], d2 = bt[2] - bvel[2];
            TV diff = (d0 < 0 ? -d0 : d0) + (d1 < 0 ? -d1 : d1) + (d2 < 0 ? -d2 : d2);
            r += -0.001 * diff;
        }
        if (collision) { dn = 1; ct[e] = 0; }
        if (ct[e] == nt) { dn = 1; ct[e] = 0; }
        if (fc) { dn = 1; ct[e] = 0; }
        rew[e] = r;
        done[e] = (unsigned char)dn;
        if (fail) fail[e] = fc;
        if (power) power[e] = (double)pwr;
    }
}

/* Raw integrator only: `substeps` calls of _run_internal (quadrotorsim.py:295-304) */
void QO_NAME(sim_step)(const qo_cfg *c, int n, double *state, const float *act, int substeps, double *power,
                       int *fail)
{
    for (int e = 0; e < n; ++e) {
        double *s = state + 22 * (long)e;
        T p[3], w[4], R[9], Ri[9];
        TV v[3], om[3];
        for (int k = 0; k < 3; ++k) { p[k] = (T)s[k]; v[k] = (TV)s[3 + k]; om[k] = (TV)s[6 + k]; }
        for (int k = 0; k < 4; ++k) w[k] = (T)s[9 + k];
        for (int k = 0; k < 9; ++k) R[k] = (T)s[13 + k];
        QO_NAME(invert)(R, Ri);
        T vclamp[4];
        for (int k = 0; k < 4; ++k) {
            double a = (double)act[4 * (long)e + k];
            if (a > c->vmax) a = c->vmax; else if (a < c->vmin) a = c->vmin;
            vclamp[k] = KT(a);
        }
        T pwr = KT(0.0);
        int fc = 0;
        for (int k = 0; k < substeps; ++k) {
            QO_NAME(substep)(c, vclamp, p, v, om, w, R, Ri, &pwr);
            fc = QO_NAME(fail)(c, p, v, om);
            if (fc) break;
        }
        for (int k = 0; k < 3; ++k) { s[k] = (double)p[k]; s[3 + k] = (double)v[k]; s[6 + k] = (double)om[k]; }
        for (int k = 0; k < 4; ++k) s[9 + k] = (double)w[k];
        for (int k = 0; k < 9; ++k) s[13 + k] = (double)R[k];
        if (power) power[e] = (double)pwr;
        if (fail) fail[e] = fc;
    }
}


/* ---------------------------------------------------------------------------------------------------------------
 * Classical RK4 on the continuous-time model behind _run_internal.  NOT a reference mode (the reference integrates
 * with semi-implicit Euler substeps only): "parity unpinned".  It exists so that the engine's RK4 option can be
 * checked (a) against this restatement and (b) by convergence of this restatement to the Euler-substep oracle as
 * precision -> 0.  All arithmetic in T.
 * --------------------------------------------------------------------------------------------------------------- */
static void QO_NAME(rhs)(const qo_cfg *c, const T vclamp[4], const T y[22], T d[22])
{
    const T *p = y, *v = y + 3, *om = y + 6, *w = y + 9, *R = y + 13;
    T Ri[9];
    (void)p;
    QO_NAME(invert)(R, Ri);
    T bv[3];
    for (int r = 0; r < 3; ++r) bv[r] = Ri[3 * r] * v[0] + Ri[3 * r + 1] * v[1] + Ri[3 * r + 2] * v[2];
    T me[4], fz = 0, tq[3] = {0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        const T px = (T)c->prop[3 * i], py = (T)c->prop[3 * i + 1];
        me[i] = (T)(c->phi / c->ra) * (vclamp[i] - (T)c->phi * w[i]);
        d[9 + i] = (me[i] - (T)c->mm) / (T)c->jm;
        const T v1 = bv[2] + (om[0] * py - om[1] * px) * (T)c->lm[i];
        const T th = (T)c->ct0 * w[i] * w[i] + (T)c->ct1 * w[i] * v1 + (T)c->ct2 * v1 * (v1 < 0 ? -v1 : v1);
        fz += th;
        tq[0] += th * py;
        tq[1] += -th * px;
    }
    tq[2] += -me[0] + me[1] - me[2] + me[3];
    const T vn = (T)QO_SQRTT(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const T on = (T)QO_SQRTT(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    T F[3], Tq[3], fg[3];
    for (int r = 0; r < 3; ++r) fg[r] = Ri[3 * r + 2] * (T)(-9.8f) * (T)c->m;
    for (int r = 0; r < 3; ++r) {
        F[r] = (r == 2 ? fz : 0) + fg[r] - vn * (T)c->Df[r] * bv[r];
        Tq[r] = tq[r] - on * (T)c->Dm[r] * om[r];
    }
    Tq[0] -= fg[1] * (T)c->cg[2] - fg[2] * (T)c->cg[1];
    Tq[1] -= fg[2] * (T)c->cg[0] - fg[0] * (T)c->cg[2];
    Tq[2] -= fg[0] * (T)c->cg[1] - fg[1] * (T)c->cg[0];
    for (int r = 0; r < 3; ++r) {
        d[r] = v[r];
        d[3 + r] = (R[3 * r] * F[0] + R[3 * r + 1] * F[1] + R[3 * r + 2] * F[2]) / (T)c->m;
        d[6 + r] = (T)c->Iinv[3 * r] * Tq[0] + (T)c->Iinv[3 * r + 1] * Tq[1] + (T)c->Iinv[3 * r + 2] * Tq[2];
        d[13 + 3 * r + 0] = R[3 * r + 1] * om[2] - R[3 * r + 2] * om[1];
        d[13 + 3 * r + 1] = R[3 * r + 2] * om[0] - R[3 * r + 0] * om[2];
        d[13 + 3 * r + 2] = R[3 * r + 0] * om[1] - R[3 * r + 1] * om[0];
    }
}

void QO_NAME(rk4_step)(const qo_cfg *c, int n, double *state, const float *act, double dt, int rk4_steps)
{
    const T h = (T)(dt / rk4_steps);
    for (int e = 0; e < n; ++e) {
        double *s = state + 22 * (long)e;
        T y[22], t[22], k[22], acc[22], vclamp[4];
        for (int q = 0; q < 22; ++q) y[q] = (T)s[q];
        for (int q = 0; q < 4; ++q) {
            double a = (double)act[4 * (long)e + q];
            if (a > c->vmax) a = c->vmax; else if (a < c->vmin) a = c->vmin;
            vclamp[q] = (T)a;
        }
        for (int it = 0; it < rk4_steps; ++it) {
            QO_NAME(rhs)(c, vclamp, y, k);
            for (int q = 0; q < 22; ++q) { acc[q] = k[q]; t[q] = y[q] + (T)0.5 * h * k[q]; }
            QO_NAME(rhs)(c, vclamp, t, k);
            for (int q = 0; q < 22; ++q) { acc[q] += 2 * k[q]; t[q] = y[q] + (T)0.5 * h * k[q]; }
            QO_NAME(rhs)(c, vclamp, t, k);
            for (int q = 0; q < 22; ++q) { acc[q] += 2 * k[q]; t[q] = y[q] + h * k[q]; }
            QO_NAME(rhs)(c, vclamp, t, k);
            for (int q = 0; q < 22; ++q) y[q] = y[q] + h / 6 * (acc[q] + k[q]);
        }
        for (int q = 0; q < 22; ++q) s[q] = (double)y[q];
    }
}

#undef KT
#undef KV
