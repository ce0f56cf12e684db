/*
 * rt_oracle.h -- CPU oracle for the per-pixel render loop of keychera/raytracing-clj.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / the CPU baseline.
 *
 * It is a double-precision restatement of the reference algorithm (Clojure, JVM):
 *   main  variant  src/raytracing.clj, src/vec3a.clj, src/hittable.clj,
 *                  src/material.clj, src/hit.clj, src/ray.clj
 *   realm variant  src/realm/raytracing.clj, src/realm/vec3.clj, src/realm/rng.clj
 *   -i    variant  src/experimental/raytracing_i.clj, src/experimental/vec3i.clj
 * The reference cannot be built or run in this image (no JVM, no Clojure), so there
 * is no oracle/_ref.  PARITY PIN: the reference has no tests and no seeded output;
 * the oracle is pinned STATISTICALLY against the reference's committed renders
 * scene.ppm (main) and scene-realm.ppm (realm) -- see tests/test_oracle_golden.py.
 * Per-ray / per-stream parity with the JVM program is unpinned (the reference's RNG
 * is unseeded); it is pinned only between this oracle and the CUDA path.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RTO_LAMBERTIAN = 0, RTO_METAL = 1, RTO_DIELECTRIC = 2 };

/* variant switches */
enum {
  RTO_F_NEAR_ZERO_GUARD = 1u,  /* material.clj:17 (main only)                       */
  RTO_F_SCHLICK = 2u,          /* material.clj:30-32,42 (main only)                 */
  RTO_F_REVERSE_PRODUCT = 4u,  /* raytracing.clj:52-53 recursion order (main)       */
  RTO_F_MEAN_DIVIDE = 8u,      /* raytracing.clj:155  sum / spp ; else sum*(1/spp)  */
  RTO_F_NORMAL_SHADING = 16u,  /* raytracing_i.clj:59-73                            */
  RTO_F_QUANT_LINEAR = 32u     /* raytracing_i.clj:169-171  int(255.999*c)          */
};

typedef struct rto_scene { /* structure of arrays, in hittable-list order */
  int32_t n;
  int32_t _pad;
  const double *center_xyz; /* [3n] */
  const double *radius;     /* [n]  */
  const int32_t *material;  /* [n]  */
  const double *albedo_rgb; /* [3n] */
  const double *fuzz;       /* [n]  */
  const double *ior;        /* [n]  */
} rto_scene;

typedef struct rto_camera { /* the vectors the reference derives before its loop */
  double pixel00[3];
  double pixel_du[3];
  double pixel_dv[3];
  double center[3];
  double defocus_u[3];
  double defocus_v[3];
  double defocus_angle; /* <= 0 : rays start at center (raytracing.clj:147) */
  int32_t width;
  int32_t height;
} rto_camera;

typedef struct rto_params {
  int32_t spp;
  int32_t max_depth;
  uint64_t seed;
  uint32_t flags;
  int32_t samples_per_unit; /* summation chunk; <=0 or >=spp : one sequential sum */
} rto_params;

typedef struct rto_stats {
  uint64_t samples;
  uint64_t segments;     /* hit-anything calls */
  uint64_t sphere_tests; /* segments * n */
  uint64_t rng_blocks;   /* Philox blocks consumed */
  uint64_t hits[3];      /* by material kind */
  uint64_t seg_hist[64]; /* samples by number of segments traced (last bin = 63+) */
} rto_stats;

/* Philox4x32-10 block function: ctr[4] -> out[4]. */
void rto_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]);

/* The uniform stream both sides use: word w (0..3) of block `block` of stage
 * `stage` of sample `sample` of pixel `pixel`, mapped to (word >> 8) * 2^-24. */
double rto_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage,
                   uint32_t block, int word);

/* Render rows [row_begin,row_end) of the image.  out_linear / out_rgb8 are FULL
 * images (W*H*3, row-major, row 0 = top); only the requested rows are written.
 * threads > 1 uses the reference's partition: contiguous chunks of
 * ceil(rows/threads) rows (raytracing.clj:157-167).  Returns 0 on success. */
int rto_render(const rto_scene *scene, const rto_camera *cam, const rto_params *prm,
               int threads, int row_begin, int row_end, double *out_linear,
               uint8_t *out_rgb8, rto_stats *stats);

/* Same, over rows row_begin, row_begin+row_step, ... < row_end (a bounded, evenly spread
 * sample of a big image for the CPU-baseline timing).  The row list is cut into
 * contiguous chunks of ceil(count/threads) rows, one per thread. */
int rto_render_strided(const rto_scene *scene, const rto_camera *cam, const rto_params *prm,
                       int threads, int row_begin, int row_end, int row_step, double *out_linear,
                       uint8_t *out_rgb8, rto_stats *stats);

/* Single closest-hit query (hit-anything) for unit tests: returns index or -1. */
int rto_hit_anything(const rto_scene *scene, const double origin[3], const double dir[3],
                     double t_min, double t_max, double *t_out, double point[3],
                     double normal[3], int *front_face);

/* Quantise one linear channel like the reference (raytracing.clj:19-26 or
 * raytracing_i.clj:170 when linear != 0). */
int rto_quantise(double c, int linear);

#ifdef __cplusplus
}
#endif
#endif
