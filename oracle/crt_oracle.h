/* crt_oracle.h -- TEST INFRASTRUCTURE ONLY (see crt_oracle.c).  CPU restatement of the reference hot path. */
#ifndef CRT_ORACLE_H
#define CRT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "../include/crtb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crt_oracle_stats {
  uint64_t rays_primary, rays_shadow, rays_reflection, rays_refraction;
  /* AABB / triangle tests under the reference's visit-all traversal, split by query kind */
  uint64_t node_tests_closest, triangle_tests_closest, node_tests_shadow, triangle_tests_shadow;
  uint64_t max_query_tests; /* most tests spent on one closest-hit query */
} crt_oracle_stats;

/* RayTracer::render (RayTracer.cpp:204-298) over options->rects; rgb = H*W*3 floats, pixels outside the rects
 * are left untouched (colorBuffer persistence); hits optional.  stats are ACCUMULATED into *stats. */
int crt_oracle_render(const crtb200_scene *scene, const crtb200_camera *camera, const crtb200_options *options,
                      float *rgb, crtb200_hit *hits, crt_oracle_stats *stats, int threads);
void crt_oracle_quantize(const float *rgb, size_t n_values, uint8_t *out);
int crt_oracle_generate_rays(const crtb200_scene *scene, const crtb200_camera *camera, float *rays_out);
int crt_oracle_trace_rays(const crtb200_scene *scene, const float *rays, uint32_t n, uint32_t ray_type,
                          const float *max_distance, crtb200_hit *hits_out, uint8_t *occluded_out);


/* Skip-rule model (see the end of crt_oracle.c): the restated reference walk with the product's conservative culling
 * applied, for checking on the CPU that the culling changes no result.  mu = per-mesh margins (crt_oracle_skip_margins). */
int crt_oracle_skip_margins(const crtb200_scene *scene, float *mu_out);
int crt_oracle_render_skip_model(const crtb200_scene *scene, const crtb200_camera *camera, const crtb200_options *options,
                                 const float *mu, float *rgb, crtb200_hit *hits, crt_oracle_stats *stats, int threads);

#ifdef __cplusplus
}
#endif
#endif
