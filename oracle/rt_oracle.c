/*
 * rt_oracle.c -- CPU oracle (double precision, plain C) for the per-pixel render
 * loop of keychera/raytracing-clj.  TEST INFRASTRUCTURE ONLY -- see rt_oracle.h.
 *
 * Build with -ffp-contract=off: the JVM never fuses a*b+c, and every expression
 * below keeps the reference's evaluation order.
 *
 * Randomness: the reference draws from unseeded Math.random / Xoshiro256++
 * (vec3a.clj:71-72, realm/rng.clj:6-10).  Both call sites are replaced here by a
 * counter-based Philox4x32-10 stream keyed by (seed; pixel, sample, stage, block)
 * so that the CUDA path can be compared draw for draw:
 *   stage 0   = camera ray: block 0 = (jitter-x, jitter-y, disk0.x, disk0.y),
 *               block n>=1 = disk candidates 2n-1 (words 0,1) and 2n (words 2,3)
 *   stage s>0 = scatter at the s-th hit of the path: block n carries unit-vector
 *               candidates 2n (words 0,1) and 2n+1 (words 2,3); a candidate's x, y, z
 *               are the three 21-bit fields of its 64 bits (lo word first), each
 *               mapped to -1 + 2 * field * 2^-21; the Schlick draw is word 0 of block 0.
 * uniform = (word >> 8) * 2^-24 everywhere else.
 */
#include "rt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Philox */

void rto_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double word_to_uniform(uint32_t w) {
  return (double)(w >> 8) * (1.0 / 16777216.0);
}

double rto_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage,
                   uint32_t block, int word) {
  uint32_t ctr[4] = {pixel, sample, stage, block}, out[4];
  rto_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), out);
  return word_to_uniform(out[word & 3]);
}

/* -------------------------------------------------------------- vec3 algebra */
/* vec3a.clj:8-69 / realm/vec3.clj:48-105.  Dot and length-squared sum left to
 * right, (x*x + y*y) + z*z; division is division. */

typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mulv(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_muls(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_divs(v3 a, double s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double v3_lensq(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline v3 v3_load(const double *p) { return v3_make(p[0], p[1], p[2]); }

/* Math/min as the JVM defines it for the one call site that matters
 * (material.clj:39, vec3a.clj:98): NaN propagates. */
static inline double jmin1(double x) { return (x != x) ? x : (x < 1.0 ? x : 1.0); }

/* --------------------------------------------------------------- the tracer */

typedef struct {
  const rto_scene *scene;
  const rto_camera *cam;
  const rto_params *prm;
  uint32_t k0, k1;
  rto_stats st;
  int32_t *att_stack; /* sphere index per scatter, for the main product order */
} tracer;

typedef struct {
  uint32_t pixel, sample, stage;
} rng_key;

static inline void draw_words(tracer *tr, rng_key k, uint32_t block, uint32_t out[4]) {
  uint32_t ctr[4] = {k.pixel, k.sample, k.stage, block};
  rto_philox4x32_10(ctr, tr->k0, tr->k1, out);
  tr->st.rng_blocks++;
}

static inline void draw_block(tracer *tr, rng_key k, uint32_t block, double u[4]) {
  uint32_t out[4];
  draw_words(tr, k, block, out);
  for (int w = 0; w < 4; ++w) u[w] = word_to_uniform(out[w]);
}

/* vec3a/rand-double -1 1 : vmin + (vmax - vmin) * rand  (vec3a.clj:71-72) */
static inline double sym(double u) { return -1.0 + 2.0 * u; }

/* vec3a/random-unit-vec3 (vec3a.clj:74-79), Realm.randUnitVec3 (realm/vec3.clj:113-121) */
static v3 random_unit(tracer *tr, rng_key k) {
  for (uint32_t block = 0;; ++block) {
    uint32_t w[4];
    draw_words(tr, k, block, w);
    for (int half = 0; half < 2; ++half) { /* two candidates per block, 3 x 21 bits each */
      uint64_t bits = (uint64_t)w[2 * half] | ((uint64_t)w[2 * half + 1] << 32);
      double x = sym((double)(bits & 0x1FFFFFu) * (1.0 / 2097152.0));
      double y = sym((double)((bits >> 21) & 0x1FFFFFu) * (1.0 / 2097152.0));
      double z = sym((double)((bits >> 42) & 0x1FFFFFu) * (1.0 / 2097152.0));
      double lensq = x * x + y * y + z * z;
      if (lensq > 1e-160 && lensq <= 1.0) return v3_divs(v3_make(x, y, z), sqrt(lensq));
    }
  }
}

typedef struct {
  int index;
  double t;
  v3 point, normal;
  int front_face;
} hit_rec;

/* hittable/sphere hit-fn (hittable.clj:9-31) = Sphere.hit (realm/raytracing.clj:96-122);
 * returns 1 and the root when the sphere is hit inside (t_min, t_max). */
static inline int sphere_root(v3 center, double radius, v3 origin, v3 dir, double t_min,
                              double t_max, double *root_out) {
  v3 oc = v3_sub(center, origin);
  double a = v3_lensq(dir);
  double h = v3_dot(dir, oc);
  double c = v3_lensq(oc) - radius * radius;
  double disc = h * h - a * c;
  if (disc < 0.0) return 0;
  double sqrt_d = sqrt(disc);
  double root = (h - sqrt_d) / a;
  if (root <= t_min || t_max <= root) {
    root = (h + sqrt_d) / a;
    if (root <= t_min || t_max <= root) return 0;
  }
  *root_out = root;
  return 1;
}

/* hit-anything (raytracing.clj:33-43) = Ray.hitAnything (realm/raytracing.clj:192-203):
 * list order, running closest-so-far as t_max; strict bounds, so the first body
 * wins an exact tie.  Point / normal / front-face as hittable.clj:24-31. */
static int hit_anything(const rto_scene *sc, v3 origin, v3 dir, double t_min, double t_max,
                        hit_rec *rec) {
  int best = -1;
  double closest = t_max;
  for (int i = 0; i < sc->n; ++i) {
    double root;
    if (sphere_root(v3_load(sc->center_xyz + 3 * i), sc->radius[i], origin, dir, t_min,
                    closest, &root)) {
      closest = root;
      best = i;
    }
  }
  if (best < 0) return 0;
  v3 center = v3_load(sc->center_xyz + 3 * best);
  rec->index = best;
  rec->t = closest;
  rec->point = v3_add(origin, v3_muls(dir, closest));              /* ray.clj:7-8 */
  v3 outward = v3_divs(v3_sub(rec->point, center), sc->radius[best]); /* hittable.clj:25 */
  rec->front_face = v3_dot(dir, outward) < 0.0;                   /* hit.clj:14-15 */
  rec->normal = rec->front_face ? outward : v3_neg(outward);
  return 1;
}

int rto_hit_anything(const rto_scene *scene, const double origin[3], const double dir[3],
                     double t_min, double t_max, double *t_out, double point[3],
                     double normal[3], int *front_face) {
  hit_rec rec;
  if (!hit_anything(scene, v3_load(origin), v3_load(dir), t_min, t_max, &rec)) return -1;
  if (t_out) *t_out = rec.t;
  if (point) { point[0] = rec.point.x; point[1] = rec.point.y; point[2] = rec.point.z; }
  if (normal) { normal[0] = rec.normal.x; normal[1] = rec.normal.y; normal[2] = rec.normal.z; }
  if (front_face) *front_face = rec.front_face;
  return rec.index;
}

/* sky gradient, raytracing.clj:55-58 / realm/raytracing.clj:229-235 */
static inline v3 sky(v3 dir) {
  double y = dir.y / sqrt(v3_lensq(dir));
  double a = 0.5 * (y + 1.0);
  return v3_make((1.0 - a) * 1.0 + a * 0.5, (1.0 - a) * 1.0 + a * 0.7,
                 (1.0 - a) * 1.0 + a * 1.0);
}

/* vec3a/reflect (vec3a.clj:94-95), Realm.reflect (realm/vec3.clj:128-133) */
static inline v3 reflect(v3 v, v3 n) { return v3_sub(v, v3_muls(n, 2.0 * v3_dot(v, n))); }

/* vec3a/refract (vec3a.clj:97-101), Realm.refract (realm/vec3.clj:135-154) */
static inline v3 refract(v3 uv, v3 n, double eta) {
  double cos_theta = jmin1(v3_dot(v3_neg(uv), n));
  v3 perp = v3_muls(v3_add(uv, v3_muls(n, cos_theta)), eta);
  v3 para = v3_muls(n, -sqrt(fabs(1.0 - v3_lensq(perp))));
  return v3_add(perp, para);
}

/* material/reflectance (material.clj:30-32).  Math/pow(x,2) and Math/pow(x,5) are
 * restated as fixed multiplication chains; the JVM's pow is a <=1-ulp function, so
 * this bit is unpinned against the real reference. */
static inline double reflectance(double cosine, double ri) {
  double q = (1.0 - ri) / (1.0 + ri);
  double r0 = q * q;
  double m = 1.0 - cosine;
  double m2 = m * m;
  double m5 = m2 * m2 * m;
  return r0 + (1.0 - r0) * m5;
}

/* One sample: camera ray (raytracing.clj:144-151 / realm :332-339) then ray-color
 * (raytracing.clj:45-58 recursive, realm/raytracing.clj:205-236 iterative). */
static v3 trace_sample(tracer *tr, uint32_t pixel, int i, int j, uint32_t sample) {
  const rto_camera *cam = tr->cam;
  const rto_scene *sc = tr->scene;
  const uint32_t flags = tr->prm->flags;
  rng_key key = {pixel, sample, 0};
  double u[4];
  draw_block(tr, key, 0, u);
  double sx = (double)i + (u[0] - 0.5);
  double sy = (double)j + (u[1] - 0.5);
  v3 p00 = v3_load(cam->pixel00), du = v3_load(cam->pixel_du), dv = v3_load(cam->pixel_dv);
  v3 pixel_sample = v3_add(v3_add(p00, v3_muls(du, sx)), v3_muls(dv, sy));
  v3 origin = v3_load(cam->center);
  if (!(cam->defocus_angle <= 0.0)) {
    /* vec3a/random-in-unit-disk (vec3a.clj:81-86) + defocus-disk-sample (raytracing.clj:89-93) */
    double px = sym(u[2]), py = sym(u[3]);
    uint32_t block = 0;
    int half = 1;
    while (!(px * px + py * py < 1.0)) {
      if (half == 1) { draw_block(tr, key, ++block, u); half = 0; } else half = 1;
      px = sym(u[2 * half]);
      py = sym(u[2 * half + 1]);
    }
    origin = v3_add(v3_add(origin, v3_muls(v3_load(cam->defocus_u), px)),
                    v3_muls(v3_load(cam->defocus_v), py));
  }
  v3 dir = v3_sub(pixel_sample, origin);

  int depth = tr->prm->max_depth;
  int nseg = 0, nstack = 0;
  v3 throughput = v3_make(1.0, 1.0, 1.0); /* realm/raytracing.clj:206 */
  v3 color = v3_make(0.0, 0.0, 0.0);
  for (;;) {
    if (depth <= 0) break; /* black: raytracing.clj:46-47, realm :209-210 */
    hit_rec rec;
    nseg++;
    if (!hit_anything(sc, origin, dir, 1e-3, INFINITY, &rec)) {
      v3 s = sky(dir);
      if (flags & RTO_F_REVERSE_PRODUCT) {
        color = s; /* ((sky * att_n) * att_{n-1}) ... * att_1, raytracing.clj:52-53 */
        for (int q = nstack - 1; q >= 0; --q) {
          int b = tr->att_stack[q];
          v3 att = sc->material[b] == RTO_DIELECTRIC ? v3_make(1.0, 1.0, 1.0)
                                                      : v3_load(sc->albedo_rgb + 3 * b);
          color = v3_mulv(color, att);
        }
      } else {
        color = v3_mulv(throughput, s); /* realm :236 */
      }
      break;
    }
    if (flags & RTO_F_NORMAL_SHADING) { /* raytracing_i.clj:62-66 */
      color = v3_muls(v3_add(rec.normal, v3_make(1.0, 1.0, 1.0)), 0.5);
      break;
    }
    key.stage++;
    int kind = sc->material[rec.index];
    tr->st.hits[kind]++;
    v3 att;
    if (kind == RTO_LAMBERTIAN) { /* material.clj:13-19, realm :138-145 */
      v3 s = v3_add(random_unit(tr, key), rec.normal);
      if ((flags & RTO_F_NEAR_ZERO_GUARD) && fabs(s.x) < 1e-8 && fabs(s.y) < 1e-8 &&
          fabs(s.z) < 1e-8)
        s = rec.normal;
      dir = s;
      att = v3_load(sc->albedo_rgb + 3 * rec.index);
    } else if (kind == RTO_METAL) { /* material.clj:21-28, realm :147-158 */
      v3 refl = reflect(dir, rec.normal);
      refl = v3_add(v3_muls(random_unit(tr, key), sc->fuzz[rec.index]), refl);
      if (!(v3_dot(refl, rec.normal) > 0.0)) break; /* absorbed: black */
      dir = refl;
      att = v3_load(sc->albedo_rgb + 3 * rec.index);
    } else { /* material.clj:34-46, realm :160-177 */
      double ior = sc->ior[rec.index];
      double ri = rec.front_face ? 1.0 / ior : ior;
      v3 unit = v3_divs(dir, sqrt(v3_lensq(dir)));
      double cos_theta = jmin1(v3_dot(v3_neg(unit), rec.normal));
      double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
      int cannot_refract = ri * sin_theta > 1.0;
      int do_reflect = cannot_refract;
      if (!do_reflect && (flags & RTO_F_SCHLICK)) { /* `or` short-circuits: material.clj:42 */
        draw_block(tr, key, 0, u);
        do_reflect = reflectance(cos_theta, ri) > u[0];
      }
      dir = do_reflect ? reflect(unit, rec.normal) : refract(unit, rec.normal, ri);
      att = v3_make(1.0, 1.0, 1.0);
    }
    origin = rec.point;
    if (flags & RTO_F_REVERSE_PRODUCT) tr->att_stack[nstack++] = rec.index;
    else throughput = v3_mulv(throughput, att); /* realm :225 */
    depth--;
  }
  tr->st.samples++;
  tr->st.segments += (uint64_t)nseg;
  tr->st.seg_hist[nseg < 63 ? nseg : 63]++;
  return color;
}

int rto_quantise(double c, int linear) {
  if (linear) { /* raytracing_i.clj:170 */
    double v = 255.999 * c;
    return (v != v) ? 0 : (int)v;
  }
  double g = c > 0.0 ? sqrt(c) : 0.0; /* linear->gamma, raytracing.clj:21-22 */
  double lo = g > 0.0 ? g : 0.0;      /* clamp, raytracing.clj:19 */
  double cl = lo < 0.999 ? lo : 0.999;
  double v = 256.0 * cl;
  return (v != v) ? 0 : (int)v;
}

static void render_rows(tracer *tr, int row_begin, int row_end, int row_step, double *out_linear,
                        uint8_t *out_rgb8) {
  const rto_camera *cam = tr->cam;
  const rto_params *prm = tr->prm;
  const int W = cam->width, spp = prm->spp;
  int unit = prm->samples_per_unit;
  if (unit <= 0 || unit > spp) unit = spp;
  const double scale = 1.0 / (double)spp; /* realm/raytracing.clj:25 */
  for (int j = row_begin; j < row_end; j += row_step) {
    for (int i = 0; i < W; ++i) {
      uint32_t pixel = (uint32_t)j * (uint32_t)W + (uint32_t)i;
      v3 acc = v3_make(0.0, 0.0, 0.0);
      for (int k0 = 0; k0 < spp; k0 += unit) {
        int k1 = k0 + unit < spp ? k0 + unit : spp;
        v3 part = v3_make(0.0, 0.0, 0.0);
        for (int k = k0; k < k1; ++k) part = v3_add(part, trace_sample(tr, pixel, i, j, (uint32_t)k));
        acc = v3_add(acc, part);
      }
      v3 px = (prm->flags & RTO_F_MEAN_DIVIDE) ? v3_divs(acc, (double)spp) : v3_muls(acc, scale);
      size_t o = 3 * (size_t)pixel;
      if (out_linear) { out_linear[o] = px.x; out_linear[o + 1] = px.y; out_linear[o + 2] = px.z; }
      if (out_rgb8) {
        int lin = (prm->flags & RTO_F_QUANT_LINEAR) != 0;
        out_rgb8[o] = (uint8_t)rto_quantise(px.x, lin);
        out_rgb8[o + 1] = (uint8_t)rto_quantise(px.y, lin);
        out_rgb8[o + 2] = (uint8_t)rto_quantise(px.z, lin);
      }
    }
  }
}

typedef struct {
  tracer tr;
  int row_begin, row_end, row_step;
  double *out_linear;
  uint8_t *out_rgb8;
} job;

static void *job_main(void *p) {
  job *jb = (job *)p;
  render_rows(&jb->tr, jb->row_begin, jb->row_end, jb->row_step, jb->out_linear, jb->out_rgb8);
  return NULL;
}

int rto_render(const rto_scene *scene, const rto_camera *cam, const rto_params *prm,
               int threads, int row_begin, int row_end, double *out_linear,
               uint8_t *out_rgb8, rto_stats *stats) {
  return rto_render_strided(scene, cam, prm, threads, row_begin, row_end, 1, out_linear, out_rgb8,
                            stats);
}

int rto_render_strided(const rto_scene *scene, const rto_camera *cam, const rto_params *prm,
                       int threads, int row_begin, int row_end, int row_step, double *out_linear,
                       uint8_t *out_rgb8, rto_stats *stats) {
  if (!scene || !cam || !prm || scene->n < 0 || cam->width <= 0 || cam->height <= 0 ||
      prm->spp <= 0)
    return 1;
  if (row_begin < 0) row_begin = 0;
  if (row_end > cam->height) row_end = cam->height;
  if (row_step < 1) row_step = 1;
  int rows = row_end > row_begin ? (row_end - row_begin + row_step - 1) / row_step : 0;
  if (rows <= 0) return 0;
  if (threads < 1) threads = 1;
  if (threads > rows) threads = rows;
  int chunk = (rows + threads - 1) / threads; /* ceil(H / pool), raytracing.clj:160 */
  job *jobs = (job *)calloc((size_t)threads, sizeof(job));
  pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  int depth_cap = prm->max_depth > 0 ? prm->max_depth : 1;
  int njobs = 0;
  for (int t = 0; t < threads; ++t) {
    int b = row_begin + t * chunk * row_step;
    int e = b + chunk * row_step < row_end ? b + chunk * row_step : row_end;
    if (b >= e) break;
    job *jb = &jobs[njobs++];
    jb->tr.scene = scene; jb->tr.cam = cam; jb->tr.prm = prm;
    jb->tr.k0 = (uint32_t)prm->seed; jb->tr.k1 = (uint32_t)(prm->seed >> 32);
    jb->tr.att_stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)depth_cap);
    jb->row_begin = b; jb->row_end = e; jb->row_step = row_step;
    jb->out_linear = out_linear; jb->out_rgb8 = out_rgb8;
  }
  if (njobs == 1) job_main(&jobs[0]);
  else {
    for (int t = 0; t < njobs; ++t) pthread_create(&tids[t], NULL, job_main, &jobs[t]);
    for (int t = 0; t < njobs; ++t) pthread_join(tids[t], NULL);
  }
  if (stats) memset(stats, 0, sizeof(*stats));
  for (int t = 0; t < njobs; ++t) {
    if (stats) {
      const rto_stats *s = &jobs[t].tr.st;
      stats->samples += s->samples; stats->segments += s->segments;
      stats->rng_blocks += s->rng_blocks;
      for (int q = 0; q < 3; ++q) stats->hits[q] += s->hits[q];
      for (int q = 0; q < 64; ++q) stats->seg_hist[q] += s->seg_hist[q];
    }
    free(jobs[t].tr.att_stack);
  }
  if (stats) stats->sphere_tests = stats->segments * (uint64_t)scene->n;
  free(jobs);
  free(tids);
  return 0;
}
