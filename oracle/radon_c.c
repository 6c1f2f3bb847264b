/* C restatement of the CT oracle (oracle/radon.py), OpenMP over rays / pixels.
 * TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/__init__.py): the reference mount
 * (/root/reference/README.md:1-5) holds no operator code; this follows the [RECALL] torch_radon v1
 * conventions spelled out in oracle/radon.py, function for function:
 *   pduo_ray_setup      <-> radon.py::ray_setup_f32        (IEEE float32, no contraction)
 *   pduo_radon_forward  <-> radon.py::radon_forward        (float64 sampling and sums)
 *   pduo_radon_backproj <-> radon.py::radon_backprojection (float64)
 *   pduo_filter         <-> radon.py::filter_sinogram      (as the Toeplitz sum over filter_taps)
 * It exists so that (a) full-size parity checks finish in seconds and (b) bench.py has a host
 * baseline that uses every core.  Built by oracle/c_port.py with
 *   gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC
 * (-ffp-contract=off keeps the float32 set-up free of fused multiply-adds, like the __f*_rn
 * intrinsics of pd_unet_b200/csrc/radon_common.cuh).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

typedef struct {
    int32_t geom, n, n_angles, det_count;
    float det_spacing, s_dist, d_dist;
    int32_t clip_to_circle;
} pduo_geom;

typedef struct {
    float xc0, yc0, vx, vy, step;
    int n_steps;
} pduo_ray;

static float guard(float d) { return d >= 0.f ? fmaxf(d, 1e-6f) : fminf(d, -1e-6f); }

static pduo_ray ray_setup(const pduo_geom* g, float cs, float sn, int d) {
    pduo_ray r;
    volatile float t1, t2;   /* volatile temporaries pin each rounding to float32 */
    const float v = (float)g->n * 0.5f;
    t1 = (float)d - (float)g->det_count * 0.5f;
    t2 = t1 + 0.5f;
    const float u = t2 * g->det_spacing;
    float sx, sy, ex, ey;
    if (g->geom == 0) { sx = u; sy = (float)g->n; ex = u; ey = -(float)g->n; }
    else { sx = 0.f; sy = g->s_dist; ex = u; ey = -g->d_dist; }
    t1 = sx * cs; t2 = sy * sn; const float rsx = t1 - t2;
    t1 = sx * sn; t2 = sy * cs; const float rsy = t1 + t2;
    t1 = ex * cs; t2 = ey * sn; const float rex = t1 - t2;
    t1 = ex * sn; t2 = ey * cs; const float rey = t1 + t2;
    t1 = rex - rsx; const float dx = guard(t1);
    t1 = rey - rsy; const float dy = guard(t1);
    float a_s, a_e;
    int hit = 1;
    if (!g->clip_to_circle) {
        t1 = -v - rsx; const float ax0 = t1 / dx;
        t1 = v - rsx;  const float ax1 = t1 / dx;
        t1 = -v - rsy; const float ay0 = t1 / dy;
        t1 = v - rsy;  const float ay1 = t1 / dy;
        a_s = fmaxf(fminf(ax0, ax1), fminf(ay0, ay1));
        a_e = fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1));
    } else {
        t1 = dx * dx; t2 = dy * dy; const float a = t1 + t2;
        t1 = rsx * dx; t2 = rsy * dy; const float b = t1 + t2;
        t1 = rsx * rsx; t2 = rsy * rsy; t1 = t1 + t2; t2 = v * v; const float c = t1 - t2;
        t1 = b * b; t2 = a * c; const float delta = t1 - t2;
        hit = delta > 0.f;
        const float sq = sqrtf(hit ? delta : 0.f);
        t1 = -b - sq; a_s = t1 / a;
        t1 = -b + sq; a_e = t1 / a;
    }
    a_s = fmaxf(a_s, 0.f);
    a_e = fminf(a_e, 1.f);
    hit = hit && (a_s < a_e);
    t1 = dx * a_s; t1 = rsx + t1; const float x0 = t1 + v;
    t1 = dy * a_s; t1 = rsy + t1; const float y0 = t1 + v;
    t1 = dx * a_e; t1 = rsx + t1; const float x1 = t1 + v;
    t1 = dy * a_e; t1 = rsy + t1; const float y1 = t1 + v;
    const float lx = x1 - x0, ly = y1 - y0;
    t1 = lx * lx; t2 = ly * ly; t1 = t1 + t2;
    const float len = sqrtf(t1);
    const int n = (int)ceilf(len);
    hit = hit && (n > 0);
    if (!hit) {
        r.xc0 = r.yc0 = r.vx = r.vy = r.step = 0.f;
        r.n_steps = -1;
        return r;
    }
    const float nf = (float)n;
    r.vx = lx / nf;
    r.vy = ly / nf;
    t1 = r.vx * r.vx; t2 = r.vy * r.vy; t1 = t1 + t2;
    r.step = sqrtf(t1);
    r.xc0 = x0 - 0.5f;
    r.yc0 = y0 - 0.5f;
    r.n_steps = n;
    return r;
}

/* out: [A, D] arrays, for the equality test against ray_setup_f32 */
void pduo_ray_setup(const pduo_geom* g, const float* trig, float* xc0, float* yc0, float* vx, float* vy, float* step,
                    int32_t* n_steps) {
    for (int a = 0; a < g->n_angles; ++a)
        for (int d = 0; d < g->det_count; ++d) {
            const pduo_ray r = ray_setup(g, trig[2 * a], trig[2 * a + 1], d);
            const long i = (long)a * g->det_count + d;
            xc0[i] = r.xc0; yc0[i] = r.yc0; vx[i] = r.vx; vy[i] = r.vy; step[i] = r.step; n_steps[i] = r.n_steps;
        }
}

static inline double pix(const double* img, int n, long ix, long iy) {
    return (ix >= 0 && ix < n && iy >= 0 && iy < n) ? img[iy * n + ix] : 0.0;
}

/* img [B, n, n] float64 -> sino [B, A, D] float64 */
void pduo_radon_forward(const double* img, double* sino, const float* trig, int batch, const pduo_geom* g) {
    const int n = g->n, A = g->n_angles, D = g->det_count;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int a = 0; a < A; ++a)
        for (int d = 0; d < D; ++d) {
            const pduo_ray r = ray_setup(g, trig[2 * a], trig[2 * a + 1], d);
            for (int b = 0; b < batch; ++b) {
                const double* im = img + (long)b * n * n;
                double acc = 0.0;
                for (int j = 0; j <= r.n_steps; ++j) {
                    const double xc = (double)r.xc0 + (double)j * (double)r.vx;
                    const double yc = (double)r.yc0 + (double)j * (double)r.vy;
                    const double xf = floor(xc), yf = floor(yc);
                    const double fx = xc - xf, fy = yc - yf;
                    const long ix = (long)xf, iy = (long)yf;
                    acc += pix(im, n, ix, iy) * ((1.0 - fy) * (1.0 - fx)) + pix(im, n, ix + 1, iy) * ((1.0 - fy) * fx) +
                           pix(im, n, ix, iy + 1) * (fy * (1.0 - fx)) + pix(im, n, ix + 1, iy + 1) * (fy * fx);
                }
                sino[((long)b * A + a) * D + d] = acc * (double)r.step;
            }
        }
}

/* sino [B, A, D] float64 -> img [B, n, n] float64 */
void pduo_radon_backproj(const double* sino, double* img, const float* trig, int batch, const pduo_geom* g) {
    const int n = g->n, A = g->n_angles, D = g->det_count;
    const double ids = 1.0 / (double)g->det_spacing;
    const double cr = D / 2.0;
    const double sd = (double)g->s_dist, k = (double)g->s_dist + (double)g->d_dist;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < batch; ++b)
        for (int y = 0; y < n; ++y) {
            const double dy = y - n / 2.0 + 0.5;
            const double* sb = sino + (long)b * A * D;
            for (int x = 0; x < n; ++x) {
                const double dx = x - n / 2.0 + 0.5;
                double acc = 0.0;
                for (int a = 0; a < A; ++a) {
                    const double cs = trig[2 * a], sn = trig[2 * a + 1];
                    double jc, w;
                    if (g->geom == 0) {
                        jc = (cs * dx + sn * dy) * ids + cr;
                        w = 1.0;
                    } else {
                        const double iden = k / (sd + sn * dx - cs * dy);
                        jc = (cs * dx + sn * dy) * ids * iden + cr;
                        w = iden;
                    }
                    const double jb = jc - 0.5;
                    const double i0f = floor(jb);
                    const double fr = jb - i0f;
                    const long i0 = (long)i0f;
                    const double* row = sb + (long)a * D;
                    const double s0 = (i0 >= 0 && i0 < D) ? row[i0] : 0.0;
                    const double s1 = (i0 + 1 >= 0 && i0 + 1 < D) ? row[i0 + 1] : 0.0;
                    acc += (s0 * (1.0 - fr) + s1 * fr) * w;
                }
                if (g->clip_to_circle && dx * dx + dy * dy > (n / 2.0) * (n / 2.0)) acc = 0.0;
                img[((long)b * n + y) * n + x] = acc * ids;
            }
        }
}

/* out[r, i] = sum_j sino[r, j] taps[(i - j) + D - 1] */
void pduo_filter(const double* sino, double* out, const double* taps, long rows, int D) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        const double* s = sino + r * D;
        for (int i = 0; i < D; ++i) {
            double acc = 0.0;
            for (int j = 0; j < D; ++j) acc += s[j] * taps[(i - j) + D - 1];
            out[r * D + i] = acc;
        }
    }
}

/* thread control for callers whose launcher pinned OMP_NUM_THREADS (torchrun sets it to 1) */
#ifdef _OPENMP
#include <omp.h>
void pduo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int pduo_get_threads(void) { return omp_get_max_threads(); }
#else
void pduo_set_threads(int n) { (void)n; }
int pduo_get_threads(void) { return 1; }
#endif
