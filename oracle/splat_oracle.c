/*
 * splat_oracle.c -- CPU restatement of the reference rasterizer (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl cpu leg may link or
 * call this file.  The product path (splatam_b200/) never does and has no CPU fallback.
 *
 * It restates, in plain C float32, the algorithm of diff-gaussian-rasterization-w-depth
 * ("X/" = /root/reference/diff-gaussian-rasterization-w-depth.git/) for the branch SplaTAM uses
 * (colors_precomp + scales/rotations, sh_degree 0):
 *   preprocess      X/cuda_rasterizer/forward.cu:155-256  (+ auxiliary.h:41-77,139-163,
 *                   computeCov3D forward.cu:118-152, computeCov2D forward.cu:74-113)
 *   keys / sort     X/cuda_rasterizer/rasterizer_impl.cu:70-138,277-319
 *   forward blend   X/cuda_rasterizer/forward.cu:261-393
 *   backward blend  X/cuda_rasterizer/backward.cu:399-557
 *   backward geom   X/cuda_rasterizer/backward.cu:144-396
 *
 * Pinning: the reference repository ships no tests or golden vectors for this path (SURVEY.md
 * section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE EXTENSION ITSELF, generated
 * on a B200 by tests/golden/make_golden.py and committed under tests/golden/ (see DESIGN.md).
 *
 * Float notes: per-Gaussian arithmetic uses fmaf() in the exact operation order nvcc emits for the
 * reference (read from its SASS), so radii / pixel centres / depths / conics are expected bit-equal
 * to the GPU.  The blend uses libm expf (the GPU uses libdevice expf, <= 2 ulp apart), so images and
 * gradients agree to ~1e-6 relative, and a pair sitting within an ulp of a threshold may flip.
 * Gradient sums are accumulated in double (the reference uses order-nondeterministic float atomics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16

typedef struct oracle_cam {
    int32_t width, height;
    float tanfovx, tanfovy, scale_modifier;
    float bg[3];
    float view[16];
    float proj[16];
} oracle_cam;

typedef struct oracle_ctx {
    oracle_cam cam;
    int P, R, gx, gy;
    float *means, *colors, *opac, *scales, *rots;
    int32_t* radii;
    float *xy, *depth, *conic_op, *cov3d;
    uint32_t *tiles_touched, *rect; /* rect: x0,y0,x1,y1 per Gaussian */
    uint64_t* keys;
    uint32_t* list;
    uint32_t* ranges; /* [tiles][2] */
    float* final_T;
    uint32_t* n_contrib;
    int rendered;
} oracle_ctx;

/* row k of a 4x4 (reference flat indexing) applied to p: fadd(fma(z,m[k+8], fma(x,m[k], y*m[k+4])), m[k+12])
 * (auxiliary.h:58-77, operation order from the reference SASS) */
static float xform_row(const float* m, int k, float x, float y, float z) {
    return fmaf(z, m[k + 8], fmaf(x, m[k], y * m[k + 4])) + m[k + 12];
}

static uint32_t higher_msb(uint32_t n) { /* rasterizer_impl.cu:35-50 */
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
    if (n >> msb) msb++;
    return msb;
}

static int f2i_trunc(float v) { /* cvt.rzi.s32.f32: saturating, NaN -> 0 */
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

static void cov3d_from_scale_rot(const float* s3, float mod, const float* q4, float* c) { /* forward.cu:118-152 */
    const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float xz_p_ry = fmaf(r, y, xz), xz_m_ry = fmaf(-r, y, xz);
    const float yz_m_rx = fmaf(y, z, -rx), yz_p_rx = fmaf(y, z, rx);
    const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
    const float xx_p_yy = fmaf(x, x, yy), xx_p_zz = fmaf(x, x, zz), yy_p_zz = yy + zz;
    const float R00 = 1.f - (yy_p_zz + yy_p_zz), R01 = xy_m_rz + xy_m_rz, R02 = xz_p_ry + xz_p_ry;
    const float R10 = xy_p_rz + xy_p_rz, R11 = 1.f - (xx_p_zz + xx_p_zz), R12 = yz_m_rx + yz_m_rx;
    const float R20 = xz_m_ry + xz_m_ry, R21 = yz_p_rx + yz_p_rx, R22 = 1.f - (xx_p_yy + xx_p_yy);
    const float sx = s3[0] * mod, sy = s3[1] * mod, sz = s3[2] * mod;
    const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
    const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
    const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
    c[0] = fmaf(M02, M02, fmaf(M00, M00, M01 * M01));
    c[1] = fmaf(M02, M12, fmaf(M00, M10, M01 * M11));
    c[2] = fmaf(M02, M22, fmaf(M00, M20, M01 * M21));
    c[3] = fmaf(M12, M12, fmaf(M10, M10, M11 * M11));
    c[4] = fmaf(M12, M22, fmaf(M10, M20, M11 * M21));
    c[5] = fmaf(M22, M22, fmaf(M20, M20, M21 * M21));
}

static void cov2d(const float* p, float fx, float fy, float tanx, float tany, const float* v, const float* vm,
                  float* cov) { /* forward.cu:74-113 */
    const float tz = xform_row(vm, 2, p[0], p[1], p[2]);
    const float tx = xform_row(vm, 0, p[0], p[1], p[2]);
    const float ty = xform_row(vm, 1, p[0], p[1], p[2]);
    const float limx = tanx * 1.3f, limy = tany * 1.3f;
    const float cx = fminf(fmaxf(tx / tz, -limx), limx), cy = fminf(fmaxf(ty / tz, -limy), limy);
    const float tz2 = tz * tz;
    const float J00 = fx / tz, J02 = ((tz * -cx) * fx) / tz2, J11 = fy / tz, J12 = ((tz * -cy) * fy) / tz2;
    const float T00 = fmaf(vm[2], J02, vm[0] * J00), T01 = fmaf(vm[6], J02, vm[4] * J00),
                T02 = fmaf(vm[10], J02, vm[8] * J00);
    const float T10 = fmaf(vm[2], J12, vm[1] * J11), T11 = fmaf(vm[6], J12, vm[5] * J11),
                T12 = fmaf(vm[10], J12, vm[9] * J11);
    const float A00 = fmaf(T02, v[2], fmaf(T00, v[0], T01 * v[1]));
    const float A10 = fmaf(T02, v[4], fmaf(T00, v[1], T01 * v[3]));
    const float A20 = fmaf(T02, v[5], fmaf(T00, v[2], T01 * v[4]));
    const float A01 = fmaf(T12, v[2], fmaf(T10, v[0], T11 * v[1]));
    const float A11 = fmaf(T12, v[4], fmaf(T10, v[1], T11 * v[3]));
    const float A21 = fmaf(T12, v[5], fmaf(T10, v[2], T11 * v[4]));
    cov[0] = fmaf(T02, A20, fmaf(T00, A00, T01 * A10)) + 0.3f;
    cov[1] = fmaf(T02, A21, fmaf(T00, A01, T01 * A11));
    cov[2] = fmaf(T12, A21, fmaf(T10, A01, T11 * A11)) + 0.3f;
}

static float ndc2pix(float v, int S) { /* auxiliary.h:41-44, evaluated in double, one DFMA */
    return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5);
}

static void preprocess(oracle_ctx* c) { /* forward.cu:155-256 */
    const oracle_cam* cam = &c->cam;
    const int W = cam->width, H = cam->height;
    const float fy = H / (2.0f * cam->tanfovy), fx = W / (2.0f * cam->tanfovx); /* rasterizer_impl.cu:222-223 */
    for (int i = 0; i < c->P; ++i) {
        c->radii[i] = 0; c->tiles_touched[i] = 0;
        c->xy[2 * i] = c->xy[2 * i + 1] = 0.f; c->depth[i] = 0.f;
        memset(c->conic_op + 4 * i, 0, 16); memset(c->rect + 4 * i, 0, 16);
        const float* p = c->means + 3 * i;
        const float depth = xform_row(cam->view, 2, p[0], p[1], p[2]);
        cov3d_from_scale_rot(c->scales + 3 * i, cam->scale_modifier, c->rots + 4 * i, c->cov3d + 6 * i);
        if (depth <= 0.2f) continue; /* auxiliary.h:154 */
        const float hx = xform_row(cam->proj, 0, p[0], p[1], p[2]);
        const float hy = xform_row(cam->proj, 1, p[0], p[1], p[2]);
        const float hw = xform_row(cam->proj, 3, p[0], p[1], p[2]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float projx = hx * pw, projy = hy * pw;
        float cov[3];
        cov2d(p, fx, fy, cam->tanfovx, cam->tanfovy, c->cov3d + 6 * i, cam->view, cov);
        const float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conx = cov[2] * det_inv, cony = cov[1] * -det_inv, conz = cov[0] * det_inv;
        const float mid = (cov[0] + cov[2]) * 0.5f;
        const float root = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf(mid + root, mid - root);
        const int radius = f2i_trunc(ceilf(sqrtf(lam) * 3.f));
        const float rf = (float)radius;
        const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
        /* getRect, auxiliary.h:46-56 */
        int x0 = f2i_trunc((px - rf) * 0.0625f), y0 = f2i_trunc((py - rf) * 0.0625f);
        int x1 = f2i_trunc((((px + rf) + 16.f) + -1.f) * 0.0625f), y1 = f2i_trunc((((py + rf) + 16.f) + -1.f) * 0.0625f);
        x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0; x1 = x1 < 0 ? 0 : x1; y1 = y1 < 0 ? 0 : y1;
        x0 = x0 > c->gx ? c->gx : x0; x1 = x1 > c->gx ? c->gx : x1;
        y0 = y0 > c->gy ? c->gy : y0; y1 = y1 > c->gy ? c->gy : y1;
        const uint32_t nt = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        if (nt == 0) continue;
        c->depth[i] = depth; c->radii[i] = radius;
        c->xy[2 * i] = px; c->xy[2 * i + 1] = py;
        c->conic_op[4 * i] = conx; c->conic_op[4 * i + 1] = cony; c->conic_op[4 * i + 2] = conz;
        c->conic_op[4 * i + 3] = c->opac[i];
        c->tiles_touched[i] = nt;
        c->rect[4 * i] = x0; c->rect[4 * i + 1] = y0; c->rect[4 * i + 2] = x1; c->rect[4 * i + 3] = y1;
    }
}

/* stable LSD radix sort of (u64 key, u32 value) over the low `bits` bits == cub::DeviceRadixSort::SortPairs
 * semantics (rasterizer_impl.cu:304-309) */
static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, size_t n, int bits) {
    if (n == 0) return;
    uint64_t* k2 = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (int shift = 0; shift < bits; shift += 8) {
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        const uint64_t mask = (bits - shift >= 8) ? 0xFFull : ((1ull << (bits - shift)) - 1);
        for (size_t i = 0; i < n; ++i) cnt[((keys[i] >> shift) & mask) + 1]++;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (size_t i = 0; i < n; ++i) { const size_t d = cnt[(keys[i] >> shift) & mask]++; k2[d] = keys[i]; v2[d] = vals[i]; }
        uint64_t* tk = keys; keys = k2; k2 = tk;
        uint32_t* tv = vals; vals = v2; v2 = tv;
    }
    /* an even number of swaps leaves the result in the caller's arrays; odd -> copy back */
    int passes = (bits + 7) / 8;
    if (passes & 1) { memcpy(k2, keys, n * sizeof(uint64_t)); memcpy(v2, vals, n * sizeof(uint32_t));
                      uint64_t* tk = keys; keys = k2; k2 = tk; uint32_t* tv = vals; vals = v2; v2 = tv; }
    free(k2); free(v2);
}

static void binning(oracle_ctx* c) { /* rasterizer_impl.cu:70-138,277-319 */
    size_t R = 0;
    for (int i = 0; i < c->P; ++i) R += c->tiles_touched[i];
    c->R = (int)R;
    c->keys = (uint64_t*)malloc((R ? R : 1) * sizeof(uint64_t));
    c->list = (uint32_t*)malloc((R ? R : 1) * sizeof(uint32_t));
    size_t off = 0;
    for (int i = 0; i < c->P; ++i) { /* duplicateWithKeys */
        if (c->radii[i] <= 0) continue;
        const uint32_t* r = c->rect + 4 * i;
        uint32_t dbits; memcpy(&dbits, &c->depth[i], 4);
        for (uint32_t y = r[1]; y < r[3]; ++y)
            for (uint32_t x = r[0]; x < r[2]; ++x) {
                c->keys[off] = ((uint64_t)(y * (uint32_t)c->gx + x) << 32) | dbits;
                c->list[off] = (uint32_t)i;
                ++off;
            }
    }
    const int bit = (int)higher_msb((uint32_t)(c->gx * c->gy));
    radix_sort_pairs(c->keys, c->list, R, 32 + bit);
    memset(c->ranges, 0, sizeof(uint32_t) * 2 * (size_t)c->gx * c->gy);
    for (size_t i = 0; i < R; ++i) { /* identifyTileRanges */
        const uint32_t t = (uint32_t)(c->keys[i] >> 32);
        if (i == 0) c->ranges[2 * t] = 0;
        else { const uint32_t pt = (uint32_t)(c->keys[i - 1] >> 32);
               if (pt != t) { c->ranges[2 * pt + 1] = (uint32_t)i; c->ranges[2 * t] = (uint32_t)i; } }
        if (i == R - 1) c->ranges[2 * t + 1] = (uint32_t)R;
    }
}

oracle_ctx* oracle_create(const oracle_cam* cam, int P, const float* means3D, const float* colors,
                          const float* opacities, const float* scales, const float* rotations) {
    oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(oracle_ctx));
    c->cam = *cam; c->P = P;
    c->gx = (cam->width + TILE - 1) / TILE; c->gy = (cam->height + TILE - 1) / TILE;
    const size_t n = P > 0 ? (size_t)P : 1, hw = (size_t)cam->width * cam->height;
#define DUP(dst, src, cnt) dst = (float*)malloc(sizeof(float) * (cnt)); memcpy(dst, src, sizeof(float) * (size_t)(P > 0 ? (cnt) : 0))
    DUP(c->means, means3D, n * 3); DUP(c->colors, colors, n * 3); DUP(c->opac, opacities, n);
    DUP(c->scales, scales, n * 3); DUP(c->rots, rotations, n * 4);
#undef DUP
    c->radii = (int32_t*)calloc(n, 4); c->xy = (float*)calloc(n * 2, 4); c->depth = (float*)calloc(n, 4);
    c->conic_op = (float*)calloc(n * 4, 4); c->cov3d = (float*)calloc(n * 6, 4);
    c->tiles_touched = (uint32_t*)calloc(n, 4); c->rect = (uint32_t*)calloc(n * 4, 4);
    c->ranges = (uint32_t*)calloc((size_t)c->gx * c->gy * 2, 4);
    c->final_T = (float*)calloc(hw, 4); c->n_contrib = (uint32_t*)calloc(hw, 4);
    preprocess(c);
    binning(c);
    return c;
}

void oracle_destroy(oracle_ctx* c) {
    if (!c) return;
    free(c->means); free(c->colors); free(c->opac); free(c->scales); free(c->rots); free(c->radii); free(c->xy);
    free(c->depth); free(c->conic_op); free(c->cov3d); free(c->tiles_touched); free(c->rect); free(c->keys);
    free(c->list); free(c->ranges); free(c->final_T); free(c->n_contrib); free(c);
}

int oracle_num_rendered(const oracle_ctx* c) { return c->R; }

void oracle_get_geometry(const oracle_ctx* c, int32_t* radii, float* xy, float* depths, float* conic_opacity,
                         uint32_t* tiles_touched, float* cov3D) {
    const size_t P = (size_t)c->P;
    if (radii) memcpy(radii, c->radii, 4 * P);
    if (xy) memcpy(xy, c->xy, 8 * P);
    if (depths) memcpy(depths, c->depth, 4 * P);
    if (conic_opacity) memcpy(conic_opacity, c->conic_op, 16 * P);
    if (tiles_touched) memcpy(tiles_touched, c->tiles_touched, 4 * P);
    if (cov3D) memcpy(cov3D, c->cov3d, 24 * P);
}

void oracle_get_binning(const oracle_ctx* c, uint64_t* keys, uint32_t* list, uint32_t* ranges) {
    if (keys) memcpy(keys, c->keys, 8 * (size_t)c->R);
    if (list) memcpy(list, c->list, 4 * (size_t)c->R);
    if (ranges) memcpy(ranges, c->ranges, 8 * (size_t)c->gx * c->gy);
}

/* pair test shared by forward and backward: returns 0 if the pair is skipped (forward.cu:336-350) */
static int pair_alpha(const float* xy, const float* con_o, float pxf, float pyf, float* dx, float* dy, float* G,
                      float* alpha) {
    *dx = xy[0] - pxf; *dy = xy[1] - pyf;
    const float sxy = fmaf(*dx, *dx * con_o[0], *dy * (*dy * con_o[2]));
    const float power = fmaf(sxy, -0.5f, -(*dy * (*dx * con_o[1])));
    if (power > 0.0f) return 0;
    *G = expf(power);
    *alpha = fminf(con_o[3] * *G, 0.99f);
    if (*alpha < 1.0f / 255.0f) return 0;
    return 1;
}

void oracle_render(oracle_ctx* c, float* out_color, float* out_depth, float* final_T, uint32_t* n_contrib) {
    /* forward.cu:261-393, one pixel at a time */
    const int W = c->cam.width, H = c->cam.height;
    const size_t hw = (size_t)W * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < c->gx * c->gy; ++t) {
        const int tx = t % c->gx, ty = t / c->gx;
        const uint32_t lo = c->ranges[2 * t], hi = c->ranges[2 * t + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                float T = 1.0f, C[3] = {0.f, 0.f, 0.f}, D = 15.0f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = lo; k < hi; ++k) {
                    contributor++;
                    const uint32_t g = c->list[k];
                    float dx, dy, G, alpha;
                    if (!pair_alpha(c->xy + 2 * g, c->conic_op + 4 * g, (float)px, (float)py, &dx, &dy, &G, &alpha)) continue;
                    const float test_T = T * (1.f - alpha);
                    if (test_T < 0.0001f) break; /* done = true (forward.cu:352-357) */
                    for (int ch = 0; ch < 3; ++ch) C[ch] = fmaf(T, alpha * c->colors[3 * g + ch], C[ch]);
                    if (T > 0.5f && test_T < 0.5f) D = c->depth[g];
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                c->final_T[pix] = T; c->n_contrib[pix] = last;
                if (final_T) final_T[pix] = T;
                if (n_contrib) n_contrib[pix] = last;
                if (out_color) for (int ch = 0; ch < 3; ++ch) out_color[ch * hw + pix] = fmaf(c->cam.bg[ch], T, C[ch]);
                if (out_depth) out_depth[pix] = D;
            }
    }
    c->rendered = 1;
}

static void geometry_backward(const oracle_ctx* c, int i, const double* acc /*9*/, float* dmeans3D, float* dscales,
                              float* drot, float* dcov_out) {
    /* computeCov2DCUDA + preprocessCUDA + computeCov3D backward, backward.cu:144-396 (float, source order) */
    const oracle_cam* cam = &c->cam;
    const float* vm = cam->view; const float* pm = cam->proj;
    const float h_y = cam->height / (2.0f * cam->tanfovy), h_x = cam->width / (2.0f * cam->tanfovx);
    const float* mean = c->means + 3 * i; const float* cov3D = c->cov3d + 6 * i;
    const float dconx = (float)acc[2], dcony = (float)acc[3], dconz = (float)acc[4];
    float t[3] = {vm[0] * mean[0] + vm[4] * mean[1] + vm[8] * mean[2] + vm[12],
                  vm[1] * mean[0] + vm[5] * mean[1] + vm[9] * mean[2] + vm[13],
                  vm[2] * mean[0] + vm[6] * mean[1] + vm[10] * mean[2] + vm[14]};
    const float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f, yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J00 = h_x / t[2], J02 = -(h_x * t[0]) / (t[2] * t[2]), J11 = h_y / t[2], J12 = -(h_y * t[1]) / (t[2] * t[2]);
    const float T00 = vm[0] * J00 + vm[2] * J02, T01 = vm[4] * J00 + vm[6] * J02, T02 = vm[8] * J00 + vm[10] * J02;
    const float T10 = vm[1] * J11 + vm[2] * J12, T11 = vm[5] * J11 + vm[6] * J12, T12 = vm[9] * J11 + vm[10] * J12;
    const float u00 = T00 * cov3D[0] + T01 * cov3D[1] + T02 * cov3D[2], u01 = T00 * cov3D[1] + T01 * cov3D[3] + T02 * cov3D[4],
                u02 = T00 * cov3D[2] + T01 * cov3D[4] + T02 * cov3D[5];
    const float u10 = T10 * cov3D[0] + T11 * cov3D[1] + T12 * cov3D[2], u11 = T10 * cov3D[1] + T11 * cov3D[3] + T12 * cov3D[4],
                u12 = T10 * cov3D[2] + T11 * cov3D[4] + T12 * cov3D[5];
    const float a = u00 * T00 + u01 * T01 + u02 * T02 + 0.3f, b = u10 * T00 + u11 * T01 + u12 * T02,
                cc = u10 * T10 + u11 * T11 + u12 * T12 + 0.3f;
    const float denom = a * cc - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0, dcov[6] = {0, 0, 0, 0, 0, 0};
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        dL_da = denom2inv * (-cc * cc * dconx + 2 * b * cc * dcony + (denom - a * cc) * dconz);
        dL_dc = denom2inv * (-a * a * dconz + 2 * a * b * dcony + (denom - a * cc) * dconx);
        dL_db = denom2inv * 2 * (b * cc * dconx - (denom + 2 * b * b) * dcony + a * b * dconz);
        dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
        dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
        dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
        dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
        dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
        dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
    }
    const float dT00 = 2 * u00 * dL_da + u10 * dL_db, dT01 = 2 * u01 * dL_da + u11 * dL_db, dT02 = 2 * u02 * dL_da + u12 * dL_db;
    const float dT10 = 2 * u10 * dL_dc + u00 * dL_db, dT11 = 2 * u11 * dL_dc + u01 * dL_db, dT12 = 2 * u12 * dL_dc + u02 * dL_db;
    const float dJ00 = vm[0] * dT00 + vm[4] * dT01 + vm[8] * dT02, dJ02 = vm[2] * dT00 + vm[6] * dT01 + vm[10] * dT02;
    const float dJ11 = vm[1] * dT10 + vm[5] * dT11 + vm[9] * dT12, dJ12 = vm[2] * dT10 + vm[6] * dT11 + vm[10] * dT12;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xg * -h_x * tz2 * dJ02, dty = yg * -h_y * tz2 * dJ12;
    const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t[0]) * tz3 * dJ02 + (2 * h_y * t[1]) * tz3 * dJ12;
    float dm[3] = {vm[0] * dtx + vm[1] * dty + vm[2] * dtz, vm[4] * dtx + vm[5] * dty + vm[6] * dtz,
                   vm[8] * dtx + vm[9] * dty + vm[10] * dtz};
    /* mean2D -> mean3D (backward.cu:366-387) */
    const float d2x = (float)acc[0], d2y = (float)acc[1];
    const float mw = 1.0f / ((pm[3] * mean[0] + pm[7] * mean[1] + pm[11] * mean[2] + pm[15]) + 0.0000001f);
    const float mul1 = (pm[0] * mean[0] + pm[4] * mean[1] + pm[8] * mean[2] + pm[12]) * mw * mw;
    const float mul2 = (pm[1] * mean[0] + pm[5] * mean[1] + pm[9] * mean[2] + pm[13]) * mw * mw;
    dm[0] += (pm[0] * mw - pm[3] * mul1) * d2x + (pm[1] * mw - pm[3] * mul2) * d2y;
    dm[1] += (pm[4] * mw - pm[7] * mul1) * d2x + (pm[5] * mw - pm[7] * mul2) * d2y;
    dm[2] += (pm[8] * mw - pm[11] * mul1) * d2x + (pm[9] * mw - pm[11] * mul2) * d2y;
    memcpy(dmeans3D + 3 * i, dm, 12);
    if (dcov_out) memcpy(dcov_out + 6 * i, dcov, 24);
    /* computeCov3D backward (backward.cu:278-341); matrices as M[c][r] */
    const float* q = c->rots + 4 * i; const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float mod = cam->scale_modifier;
    const float s[3] = {mod * c->scales[3 * i], mod * c->scales[3 * i + 1], mod * c->scales[3 * i + 2]};
    const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                            {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                            {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float M[3][3], dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                               {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}, dM[3][3], dMt[3][3];
    for (int cI = 0; cI < 3; ++cI) for (int rI = 0; rI < 3; ++rI) M[cI][rI] = s[rI] * Rm[cI][rI];
    for (int cI = 0; cI < 3; ++cI) for (int rI = 0; rI < 3; ++rI)
        dM[cI][rI] = 2.f * (M[0][rI] * dS[cI][0] + M[1][rI] * dS[cI][1] + M[2][rI] * dS[cI][2]);
    for (int k = 0; k < 3; ++k) dscales[3 * i + k] = Rm[0][k] * dM[0][k] + Rm[1][k] * dM[1][k] + Rm[2][k] * dM[2][k];
    for (int cI = 0; cI < 3; ++cI) for (int rI = 0; rI < 3; ++rI) dMt[cI][rI] = s[cI] * dM[rI][cI];
    drot[4 * i + 0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
    drot[4 * i + 1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
    drot[4 * i + 2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
    drot[4 * i + 3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
}

/* dL_dpix [3,H,W] -> the six gradients SplaTAM consumes.  Requires oracle_render() first. */
int oracle_backward(oracle_ctx* c, const float* dL_dpix, float* dmeans3D, float* dmeans2D, float* dcolors,
                    float* dopacity, float* dscales, float* drot) {
    if (!c->rendered) return -1;
    const int W = c->cam.width, H = c->cam.height, P = c->P;
    const size_t hw = (size_t)W * H;
    double* acc = (double*)calloc((size_t)(P > 0 ? P : 1) * 9, sizeof(double));
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    /* backward.cu:399-557, sequential over pixels so the double accumulation is deterministic */
    for (int t = 0; t < c->gx * c->gy; ++t) {
        const int tx = t % c->gx, ty = t / c->gx;
        const uint32_t lo = c->ranges[2 * t];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)py * W + px;
                const float T_final = c->final_T[pix];
                float T = T_final;
                const uint32_t last = c->n_contrib[pix];
                const float dL[3] = {dL_dpix[pix], dL_dpix[hw + pix], dL_dpix[2 * hw + pix]};
                float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
                for (uint32_t kk = last; kk-- > 0;) {
                    const uint32_t g = c->list[lo + kk];
                    float dx, dy, G, alpha;
                    if (!pair_alpha(c->xy + 2 * g, c->conic_op + 4 * g, (float)px, (float)py, &dx, &dy, &G, &alpha)) continue;
                    const float* con_o = c->conic_op + 4 * g;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    double* a = acc + 9 * (size_t)g;
                    for (int ch = 0; ch < 3; ++ch) {
                        const float col = c->colors[3 * g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = col;
                        dL_dalpha += (col - accum_rec[ch]) * dL[ch];
                        a[6 + ch] += (double)(dchannel_dcolor * dL[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += c->cam.bg[ch] * dL[ch];
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = con_o[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * con_o[0] - gdy * con_o[1];
                    const float dG_ddely = -gdy * con_o[2] - gdx * con_o[1];
                    a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                    a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                    a[2] += (double)(-0.5f * gdx * dx * dL_dG);
                    a[3] += (double)(-0.5f * gdx * dy * dL_dG);
                    a[4] += (double)(-0.5f * gdy * dy * dL_dG);
                    a[5] += (double)(G * dL_dalpha);
                }
            }
    }
    for (int i = 0; i < P; ++i) {
        const double* a = acc + 9 * (size_t)i;
        dmeans2D[3 * i] = (float)a[0]; dmeans2D[3 * i + 1] = (float)a[1]; dmeans2D[3 * i + 2] = 0.f;
        dopacity[i] = (float)a[5];
        for (int ch = 0; ch < 3; ++ch) dcolors[3 * i + ch] = (float)a[6 + ch];
        if (c->radii[i] > 0) geometry_backward(c, i, a, dmeans3D, dscales, drot, NULL);
        else { memset(dmeans3D + 3 * i, 0, 12); memset(dscales + 3 * i, 0, 12); memset(drot + 4 * i, 0, 16); }
    }
    free(acc);
    return 0;
}

void oracle_mark_visible(const oracle_cam* cam, int P, const float* means3D, uint8_t* present) {
    /* checkFrustum, rasterizer_impl.cu:54-66 */
    for (int i = 0; i < P; ++i)
        present[i] = (xform_row(cam->view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) <= 0.2f) ? 0 : 1;
}
