/* A plain-C client of include/rtclj_b200.h, linked against librtclj_b200.so.
 * Proves the header is valid C11 (a JNA / Panama jextract user consumes it as C), that the POD
 * struct layouts are the documented ones, and that the host-only entry points work from C.
 * With the argument "gpu" it also renders the reference's default scene at a small size through
 * rtclj_render and writes the P3 text with both writers (tests/test_gpu_parity.py compares).
 * Built and run by tests/test_abi.py. */
#include "rtclj_b200.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

_Static_assert(sizeof(rtclj_scene) == 56, "rtclj_scene layout");
_Static_assert(sizeof(rtclj_camera) == 160, "rtclj_camera layout");
_Static_assert(sizeof(rtclj_params) == 48, "rtclj_params layout");
_Static_assert(offsetof(rtclj_camera, defocus_angle) == 144, "rtclj_camera.defocus_angle");
_Static_assert(offsetof(rtclj_params, seed) == 8, "rtclj_params.seed");

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      fprintf(stderr, "%s:%d: %s -- %s\n", __FILE__, __LINE__, #cond, rtclj_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char **argv) {
  CHECK(rtclj_abi_version() == RTCLJ_ABI_VERSION);
  /* (/ 16 9) as Clojure turns it into a double: one ulp above 16.0/9.0 */
  CHECK(rtclj_ratio_to_double(16, 9) == 0x1.c71c71c71c71dp+0);
  const double from[3] = {-2, 2, 1}, at[3] = {0, 0, -1}, up[3] = {0, 1, 0};
  rtclj_camera cam;
  CHECK(rtclj_camera_main(64, 36, 20.0, from, at, up, 10.0, 3.4, &cam) == RTCLJ_OK);
  CHECK(cam.width == 64 && cam.height == 36);
  CHECK(rtclj_camera_main(0, 36, 20.0, from, at, up, 10.0, 3.4, &cam) == RTCLJ_E_INVALID);
  CHECK(rtclj_camera_main(64, 36, 20.0, from, at, up, 10.0, 3.4, &cam) == RTCLJ_OK);

  const uint8_t px[2 * 3] = {0, 9, 10, 99, 100, 255};
  char text[64];
  size_t len = 0;
  CHECK(rtclj_encode_ppm_p3(px, 2, 1, text, sizeof text, &len) == RTCLJ_OK);
  CHECK(len == strlen("P3\n2 1\n255\n0 9 10\n99 100 255\n") && memcmp(text, "P3\n2 1\n255\n0 9 10\n99 100 255\n", len) == 0);
  const double lin[3] = {0.25, 1.5, -1.0};
  uint8_t q[3];
  CHECK(rtclj_quantise_rgb8(lin, 3, 0, q) == RTCLJ_OK && q[0] == 128 && q[1] == 255 && q[2] == 0);

  if (argc < 2 || strcmp(argv[1], "gpu") != 0) { puts("ok host"); return 0; }

  /* raytracing.clj:63-78, list order */
  const double centers[5 * 3] = {0, -100.5, -1, 0, 0, -1.2, -1, 0, -1, -1, 0, -1, 1, 0, -1};
  const double radii[5] = {100, 0.5, 0.5, 0.4, 0.5};
  const int32_t kinds[5] = {0, 0, 2, 2, 1};
  const double albedo[5 * 3] = {0.8, 0.8, 0.0, 0.1, 0.2, 0.5, 1, 1, 1, 1, 1, 1, 0.8, 0.6, 0.2};
  const double fuzz[5] = {0, 0, 0, 0, 1.0};
  const double ior[5] = {1, 1, 1.5, 1.0 / 1.5, 1};
  rtclj_scene sc;
  memset(&sc, 0, sizeof sc);
  sc.n = 5; sc.center_xyz = centers; sc.radius = radii; sc.material = kinds;
  sc.albedo_rgb = albedo; sc.fuzz = fuzz; sc.ior = ior;
  rtclj_params prm;
  memset(&prm, 0, sizeof prm);
  prm.spp = 8; prm.max_depth = 50; prm.seed = 1; prm.flags = RTCLJ_FLAGS_MAIN; prm.samples_per_unit = 8;
  const size_t n = (size_t)cam.width * cam.height * 3;
  double *linear = malloc(n * sizeof *linear);
  uint8_t *rgb = malloc(n);
  rtclj_stats st;
  CHECK(rtclj_render(&sc, &cam, &prm, linear, rgb, &st) == RTCLJ_OK);
  CHECK(st.samples == (uint64_t)cam.width * cam.height * 8 && st.segments >= st.samples);
  size_t cap = 0, l1 = 0, l2 = 0;
  CHECK(rtclj_encode_ppm_p3(NULL, cam.width, cam.height, NULL, 0, &cap) == RTCLJ_OK);
  char *t1 = malloc(cap), *t2 = malloc(cap);
  CHECK(rtclj_encode_ppm_p3(rgb, cam.width, cam.height, t1, cap, &l1) == RTCLJ_OK);
  CHECK(rtclj_encode_ppm_p3_gpu(0, rgb, cam.width, cam.height, t2, cap, &l2) == RTCLJ_OK);
  CHECK(l1 == l2 && memcmp(t1, t2, l1) == 0);
  if (argc > 2) {
    FILE *f = fopen(argv[2], "wb");
    CHECK(f != NULL);
    fwrite(linear, sizeof *linear, n, f);
    fclose(f);
  }
  /* an error from the compute path is reported, not thrown */
  prm.spp = 0;
  CHECK(rtclj_render(&sc, &cam, &prm, linear, rgb, &st) == RTCLJ_E_INVALID && strlen(rtclj_last_error()) > 0);
  free(linear); free(rgb); free(t1); free(t2);
  puts("ok gpu");
  return 0;
}
