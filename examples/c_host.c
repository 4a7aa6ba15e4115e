/* A plain-C caller of the drop-in boundary (include/vermeer_gpu.h), the way a cgo binding would use it:
 *
 *     gcc -std=c99 -Iinclude -o c_host examples/c_host.c -Lvermeer_b200 -lvermeer_b200 -Wl,-rpath,$PWD/vermeer_b200
 *     ./c_host examples/cornell64.vnf
 *
 * Host layer (no GPU needed): nodes.Parse + core.PreRender of a .vnf scene, then the structures the reference would hold
 * (scene tree, per-mesh QBVHs). Device layer (skipped when no GPU is present): upload, framescramble, four iterations of the
 * Render loop, mean pixel value. Exit code 0 = every call succeeded. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "vermeer_gpu.h"

static uint64_t splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s scene.vnf\n", argv[0]);
    return 2;
  }
  vh_scene* sc = NULL;
  if (vh_scene_create(&sc) != 0) return 1;
  const int nerr = vh_load_vnf(sc, argv[1]);
  if (nerr != 0) {
    fprintf(stderr, "parse: %d error(s)\n%s\n", nerr, vh_last_error(sc));
    return 1;
  }
  if (vh_prerender(sc) != 0) {
    fprintf(stderr, "prerender: %s\n", vh_last_error(sc));
    return 1;
  }
  int32_t g[3], info[4];
  vh_globals(sc, g);
  vh_scene_info(sc, info);
  printf("host: %dx%d, MaxIter %d, %d geoms, scene tree %d nodes over %d slots\n", g[0], g[1], g[2], vh_num_geoms(sc), info[0], info[3]);

  vg_ctx* ctx = NULL;
  if (vg_create(&ctx, 0) != 0) {
    printf("device: none (%s) - host layer only\n", vg_last_error(NULL));
    vh_scene_destroy(sc);
    return 0;
  }
  if (vh_upload(sc, ctx, 0) != 0) {
    fprintf(stderr, "upload: %s\n", vh_last_error(sc));
    return 1;
  }
  const int64_t npix = (int64_t)g[0] * g[1];
  uint64_t* table = (uint64_t*)malloc((size_t)npix * 6 * sizeof(uint64_t));
  float* fb = (float*)malloc((size_t)npix * 3 * sizeof(float));
  uint64_t seed = 1;
  for (int64_t i = 0; i < npix * 6; i++) table[i] = splitmix64(&seed);
  if (vg_set_scramble(ctx, table, npix) != 0 || vg_render(ctx, 0, 4, fb) != 0) {
    fprintf(stderr, "render: %s\n", vg_last_error(ctx));
    return 1;
  }
  double sum = 0;
  int64_t n = 0;
  for (int64_t i = 0; i < npix * 3; i++)
    if (fb[i] == fb[i]) { sum += fb[i]; n++; }   /* (the reference's Cornell frame holds a few NaN pixels: DESIGN.md quirk list) */
  VgStats st;
  vg_get_stats(ctx, &st);
  printf("device: 4 iterations, %llu rays (%llu shadow), mean pixel %.6f\n", (unsigned long long)st.rays, (unsigned long long)st.shadow_rays, sum / (double)n);
  free(table);
  free(fb);
  vg_destroy(ctx);
  vh_scene_destroy(sc);
  return 0;
}
