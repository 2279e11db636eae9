/*
 * mapper_loop.c -- the mapping loop of Gaussian-LIC (mapping.cpp:124-201: per keyframe extend() then optimize(), evaluation
 * and map export at the end) on the native mapper, from plain C: no LibTorch, no Python, one shared library.
 *
 *   gcc -std=c99 -O2 -I include examples/mapper_loop.c -o mapper_loop -L gaussian_lic_b200 -l:libglic_b200.so \
 *       -Wl,-rpath,$PWD/gaussian_lic_b200 -lm
 *   ./mapper_loop [keyframes] [initial_gaussians]        (needs a CUDA device; exits 2 with the library's message otherwise)
 *
 * Synthetic inputs (a fixed LCG): Gaussians scattered in the frustum of the first camera, keyframes on a small arc looking at
 * the scene, one smooth target image per keyframe, a LiDAR sweep of 4000 points per keyframe.  Prints one line per keyframe.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "glic_b200.h"

static unsigned long long lcg = 88172645463325252ULL;
static float urand(void) {                      /* uniform in [0, 1) */
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    return (float)((lcg >> 40) & 0xFFFFFF) / 16777216.0f;
}
static float nrand(void) { return sqrtf(-2.0f * logf(urand() + 1e-12f)) * cosf(6.2831853f * urand()); }

#define CHECK(call)                                                                           \
    do {                                                                                      \
        int st_ = (call);                                                                     \
        if (st_ != GLIC_OK) {                                                                 \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, glic_last_error());         \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

int main(int argc, char** argv) {
    const int keyframes = argc > 1 ? atoi(argv[1]) : 4;
    const unsigned P0 = argc > 2 ? (unsigned)atoi(argv[2]) : 20000u;
    const int W = 320, H = 208, n_lidar = 4000;
    const float f = 250.0f;

    glic_mapper_config cfg;
    cfg.width = W; cfg.height = H; cfg.fx = f; cfg.fy = f; cfg.cx = W / 2.0f; cfg.cy = H / 2.0f;
    cfg.sh_degree = 3;
    cfg.position_lr = 1.6e-4f; cfg.feature_lr = 2.5e-3f; cfg.opacity_lr = 0.05f; cfg.scaling_lr = 0.005f; cfg.rotation_lr = 0.001f;
    cfg.lambda_dssim = 0.2f; cfg.scaling_scale = 1.0f;
    cfg.capacity = P0 / 2;                      /* deliberately too small: initialize() and extend() grow the arena */
    cfg.max_iters = 100; cfg.seed = 7; cfg.rank = 0; cfg.world = 1; cfg.views_per_rank = 1;
    glic_mapper* m = NULL;
    CHECK(glic_mapper_create(&cfg, &m));

    /* initial map: left 70 % of the first view's frustum */
    float* xyz = malloc(sizeof(float) * 3 * P0), *dc = malloc(sizeof(float) * 3 * P0), *op = malloc(sizeof(float) * P0);
    float* ls = malloc(sizeof(float) * 3 * P0), *rot = malloc(sizeof(float) * 4 * P0);
    for (unsigned i = 0; i < P0; ++i) {
        const float z = 1.0f + 11.0f * urand();
        xyz[3 * i] = z * (urand() * 1.4f - 1.0f) * W / (2 * f); xyz[3 * i + 1] = z * (urand() * 2.0f - 1.0f) * H / (2 * f); xyz[3 * i + 2] = z;
        for (int c = 0; c < 3; ++c) { dc[3 * i + c] = nrand(); ls[3 * i + c] = -3.0f + 0.6f * nrand(); }
        op[i] = 2.0f * nrand();
        float n = 0.f;
        for (int c = 0; c < 4; ++c) { rot[4 * i + c] = nrand(); n += rot[4 * i + c] * rot[4 * i + c]; }
        for (int c = 0; c < 4; ++c) rot[4 * i + c] /= sqrtf(n);
    }
    CHECK(glic_mapper_initialize(m, P0, xyz, dc, NULL, op, ls, rot));

    float* images = malloc(sizeof(float) * 3 * W * H * (size_t)keyframes);
    float* pts = malloc(sizeof(float) * 3 * n_lidar), *col = malloc(sizeof(float) * 3 * n_lidar), *dep = malloc(sizeof(float) * n_lidar);
    float first_loss = -1.f, last_loss = -1.f;
    for (int k = 0; k < keyframes; ++k) {
        /* camera on a small arc around (0, 0, 10), looking at it: R_wc columns = right, down, forward */
        const float ang = 0.08f * (float)k;
        const float pos[3] = {2.0f * sinf(ang), 0.0f, 2.0f * (1.0f - cosf(ang))};
        float fw[3] = {0.f - pos[0], 0.f - pos[1], 10.f - pos[2]};
        const float fn = sqrtf(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
        for (int c = 0; c < 3; ++c) fw[c] /= fn;
        const float rt[3] = {fw[2], 0.f, -fw[0]};                                 /* forward x (0,-1,0), normalised (forward.y = 0) */
        const float rn = sqrtf(rt[0] * rt[0] + rt[2] * rt[2]);
        const float right[3] = {rt[0] / rn, 0.f, rt[2] / rn};
        const float down[3] = {fw[1] * right[2] - fw[2] * right[1], fw[2] * right[0] - fw[0] * right[2], fw[0] * right[1] - fw[1] * right[0]};
        glic_keyframe kf;
        for (int r = 0; r < 3; ++r) { kf.R_wc[3 * r] = right[r]; kf.R_wc[3 * r + 1] = down[r]; kf.R_wc[3 * r + 2] = fw[r]; kf.t_wc[r] = pos[r]; }
        float* img = images + (size_t)k * 3 * W * H;
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    img[((size_t)c * H + y) * W + x] = 0.5f + 0.4f * sinf(0.02f * (float)(x + 7 * c) + 0.3f * (float)k) * cosf(0.03f * (float)y);
        kf.image = img;
        CHECK(glic_mapper_add_keyframe(m, &kf, 1));
        for (int i = 0; i < n_lidar; ++i) {                                       /* a sweep in this camera's frustum, world frame */
            const float z = 0.5f + 12.0f * urand();
            const float pc[3] = {z * (urand() * 2.0f - 1.0f) * W / (2 * f), z * (urand() * 2.0f - 1.0f) * H / (2 * f), z};
            for (int r = 0; r < 3; ++r) pts[3 * i + r] = kf.R_wc[3 * r] * pc[0] + kf.R_wc[3 * r + 1] * pc[1] + kf.R_wc[3 * r + 2] * pc[2] + pos[r];
            for (int c = 0; c < 3; ++c) col[3 * i + c] = urand();
            dep[i] = z;
        }
        CHECK(glic_mapper_extend(m, n_lidar, pts, col, dep));
        CHECK(glic_mapper_optimize(m, NULL, 0));
        glic_mapper_stats st;
        CHECK(glic_mapper_stats_get(m, &st));
        if (k == 0) first_loss = st.last_loss;
        last_loss = st.last_loss;
        printf("keyframe %d: inserted %u -> %u Gaussians (capacity %u), %llu iterations so far, loss %.5f, extend %.2f ms, optimize %.2f ms\n",
               k, st.last_inserted, st.num_gaussians, st.capacity, (unsigned long long)st.iterations, st.last_loss, st.ms_extend, st.ms_optimize);
    }
    for (int rep = 0; rep < 5; ++rep) CHECK(glic_mapper_optimize(m, NULL, 0));      /* a few more passes over all keyframes */
    glic_mapper_stats st;
    CHECK(glic_mapper_stats_get(m, &st));
    last_loss = st.last_loss;
    float psnr = 0.f, ssim = 0.f;
    CHECK(glic_mapper_evaluate(m, 1, keyframes - 1, &psnr, &ssim));
    CHECK(glic_mapper_save_map(m, "/tmp/glic_mapper_loop_point_cloud.ply"));
    printf("done: %u Gaussians, %llu iterations, loss %.5f -> %.5f, last keyframe PSNR %.2f dB SSIM %.4f, map at /tmp/glic_mapper_loop_point_cloud.ply\n",
           st.num_gaussians, (unsigned long long)st.iterations, first_loss, last_loss, psnr, ssim);
    CHECK(glic_mapper_destroy(m));
    free(xyz); free(dc); free(op); free(ls); free(rot); free(images); free(pts); free(col); free(dep);
    return (last_loss == last_loss && last_loss < first_loss && st.capacity >= st.num_gaussians) ? 0 : 1;
}
