/* crt_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product path
 * (libcrtb200.so / libcrtfront.so / the Python package); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it, and only as the checker.
 *
 * Plain-C restatement of the reference's per-pixel hot path, operating on the same flattened scene the C ABI takes
 * (include/crtb200.h).  Each function cites the reference file:line it follows (paths relative to
 * /root/reference/SourceCode).  Arithmetic: binary32, evaluated in the reference's expression order; build with
 * -ffp-contract=off and no -march=native (the reference's x86-64 build cannot contract FMAs either).
 *
 * PARITY PIN: the reference holds no tests, golden vectors or fixtures for this path, so this
 * restatement is pinned against the reference ITSELF, compiled unmodified into oracle/_ref/ (oracle/build_ref.sh):
 * tests/test_oracle_vs_reference.py requires byte-identical float RGB and identical hit ids on every authored scene,
 * and the outputs of that reference build are committed as fixtures under tests/golden/ (tests/golden/make_golden.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/crtb200.h"
#include "crt_oracle.h"

typedef struct { float x, y, z; } v3;

/* ---- Vector (src/Vector.cpp) ---- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }              /* Vector.cpp:34-40 */
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }              /* Vector.cpp:42-48 */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            /* Vector.cpp:57-59 */
static inline v3 vcross(v3 a, v3 b) {                                                         /* Vector.cpp:61-65 */
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }               /* Vector.cpp:67-69 */
static inline v3 sscale(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }               /* Vector.cpp:71-73 */
static inline float vlen(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }           /* Vector.cpp:114-117 */
static inline v3 vnorm(v3 a) {                                                                /* Vector.cpp:97-106 */
  float l = vlen(a);
  if (l == 0) return a;
  l = 1.0f / l;
  return V(a.x * l, a.y * l, a.z * l);
}
static inline v3 vreflect(v3 d, v3 n) { return vsub(d, sscale(2 * vdot(d, n), n)); }          /* Vector.cpp:119-122 */
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline v3 ld3(const float *p) { return V(p[0], p[1], p[2]); }

/* std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a  -- NaN operands propagate like libstdc++ */
static inline float stdmax(float a, float b) { return (a < b) ? b : a; }
static inline float stdmin(float a, float b) { return (b < a) ? b : a; }

typedef struct { v3 o, d; int type; } ray_t;

typedef struct {
  const crtb200_scene *s;
  uint32_t max_depth;
  float shadow_bias, reflection_bias, refraction_bias;
  /* per-thread counters */
  uint64_t rays[4];
  uint64_t node_tests[2], tri_tests[2]; /* [0] closest-hit queries, [1] shadow queries */
  int shadow_query;
  uint64_t max_query_tests; /* most AABB + triangle tests spent on a single query (latency tail) */
  /* skip-rule model (crt_oracle_render_skip_model, see the end of this file); all zero for the reference walk */
  const float *skip_mu;   /* per-mesh margin, NULL = the reference's visit-all walk */
  float shadow_limit;     /* distance limit of the current shadow query */
} octx;

static int skip_node(const octx *c, const crtb200_kdnode *n, const ray_t *r, float mu, float limit, int allow_behind);

/* ---- BoundingBox::hasIntersection  include/tracer/BoundingBox.h:85-108 ---- */
static int box_hit(const crtb200_kdnode *n, const ray_t *r) {
  float t0 = -FLT_MAX, t1 = FLT_MAX;
  for (int i = 0; i < 3; i++) {
    float d = comp(r->d, i), o = comp(r->o, i);
    if (fabsf(d) < FLT_EPSILON) {
      if (o < n->box_min[i] || o > n->box_max[i]) return 0;
    } else {
      float inv = 1.0f / d;
      float tn = (n->box_min[i] - o) * inv;
      float tf = (n->box_max[i] - o) * inv;
      if (tn > tf) { float t = tn; tn = tf; tf = t; }
      t0 = stdmax(t0, tn);
      t1 = stdmin(t1, tf);
      if (t0 > t1) return 0;
    }
  }
  return 1;
}

/* ---- Triangle::pointIsInTriangle  src/Triangle.cpp:37-57 ---- */
static int point_in_tri(v3 n, v3 v0, v3 v1, v3 v2, v3 p) {
  v3 e0 = vsub(v1, v0), c0 = vsub(p, v0);
  if (vdot(n, vcross(e0, c0)) < -FLT_EPSILON) return 0;
  v3 e1 = vsub(v2, v1), c1 = vsub(p, v1);
  if (vdot(n, vcross(e1, c1)) < -FLT_EPSILON) return 0;
  v3 e2 = vsub(v0, v2), c2 = vsub(p, v2);
  if (vdot(n, vcross(e2, c2)) < -FLT_EPSILON) return 0;
  return 1;
}

/* ---- Ray::intersectWithTriangle  src/Ray.cpp:9-31 ---- */
static int ray_tri(const crtb200_scene *s, const ray_t *r, uint32_t tri, float *t_out, v3 *p_out) {
  const uint32_t *iv = s->triangle_vertex + 3 * (size_t)tri;
  v3 n = ld3(s->triangle_normal + 3 * (size_t)tri);
  v3 v0 = ld3(s->vertex_position + 3 * (size_t)iv[0]);
  float nd = vdot(r->d, n);
  if (r->type == CRTB200_RAY_PRIMARY && nd >= 0) return 0;
  float dist = -vdot(v0, n);
  float t = -(vdot(n, r->o) + dist) / nd;
  if (t < 0) return 0;
  v3 p = vadd(r->o, vscale(r->d, t));
  v3 v1 = ld3(s->vertex_position + 3 * (size_t)iv[1]);
  v3 v2 = ld3(s->vertex_position + 3 * (size_t)iv[2]);
  if (!point_in_tri(n, v0, v1, v2, p)) return 0;
  *t_out = t;
  *p_out = p;
  return 1;
}

typedef struct {
  int has;
  uint32_t mesh, tri; /* tri = GLOBAL triangle index */
  float t;
  v3 p, n;
  float u, v;
} hitinfo;

#define STACK_MAX 4096

/* ---- KDTree<Triangle>::intersect  src/KDTree.cpp:48-87 ---- */
static hitinfo mesh_intersect(octx *c, uint32_t mesh_index, const ray_t *r, float outer_min_t) {
  const crtb200_scene *s = c->s;
  const crtb200_mesh *m = &s->meshes[mesh_index];
  const crtb200_kdnode *nodes = s->mesh_nodes + m->first_node;
  const uint32_t *refs = s->mesh_leaf_refs + m->first_leaf_ref;
  hitinfo best;
  memset(&best, 0, sizeof(best));
  float min_t = INFINITY;
  uint32_t stack[STACK_MAX];
  int sp = 0;
  if (m->n_nodes == 0) return best;
  stack[sp++] = 0;
  while (sp > 0) {
    const crtb200_kdnode *n = &nodes[stack[--sp]];
    c->node_tests[c->shadow_query]++;
    if (!box_hit(n, r)) continue;
    if (c->skip_mu) { /* model of the product's conservative skip; never taken by the reference walk */
      const float lim = c->shadow_query ? c->shadow_limit : (min_t < outer_min_t ? min_t : outer_min_t);
      if (skip_node(c, n, r, c->skip_mu[mesh_index], lim, c->shadow_query || lim < INFINITY)) continue;
    }
    if (n->leaf_count) {
      for (uint32_t k = 0; k < n->leaf_count; k++) {
        uint32_t tri = m->first_triangle + refs[n->leaf_start + k];
        float t;
        v3 p;
        c->tri_tests[c->shadow_query]++;
        if (ray_tri(s, r, tri, &t, &p)) {
          /* intersections[0] is the initial closest (KDTree.cpp:78); strict < afterwards (:80-85) */
          if (!best.has) {
            best.has = 1; best.mesh = mesh_index; best.tri = tri; best.t = t; best.p = p;
          }
          if (t < min_t) {
            min_t = t; best.mesh = mesh_index; best.tri = tri; best.t = t; best.p = p;
          }
        }
      }
    } else {
      if (n->child[0] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[0];
      if (n->child[1] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[1];
    }
  }
  if (best.has) best.n = ld3(s->triangle_normal + 3 * (size_t)best.tri); /* Ray.cpp:28 hitNormal = triangleNormal */
  return best;
}

/* ---- Triangle::getBarycentricCoordinates  src/Triangle.cpp:63-73 ---- */
static void barycentric(const crtb200_scene *s, uint32_t tri, v3 p, float *u, float *v) {
  const uint32_t *iv = s->triangle_vertex + 3 * (size_t)tri;
  v3 v0 = ld3(s->vertex_position + 3 * (size_t)iv[0]);
  v3 v1 = ld3(s->vertex_position + 3 * (size_t)iv[1]);
  v3 v2 = ld3(s->vertex_position + 3 * (size_t)iv[2]);
  v3 v0p = vsub(p, v0), v0v1 = vsub(v1, v0), v0v2 = vsub(v2, v0);
  float area = vlen(vcross(v0v1, v0v2));
  *u = vlen(vcross(v0p, v0v2)) / area;
  *v = vlen(vcross(v0v1, v0p)) / area;
}

/* ---- KDTree<ObjectKDTreeSubTree>::intersect  src/KDTree.cpp:127-192 ---- */
static hitinfo scene_intersect(octx *c, const ray_t *r) {
  const crtb200_scene *s = c->s;
  hitinfo best;
  memset(&best, 0, sizeof(best));
  float min_t = INFINITY;
  uint32_t stack[STACK_MAX];
  int sp = 0;
  if (r->type >= 0 && r->type < 4) c->rays[r->type]++;
  c->shadow_query = 0;
  const uint64_t before = c->node_tests[0] + c->tri_tests[0];
  if (s->n_top_nodes == 0) return best;
  stack[sp++] = 0;
  while (sp > 0) {
    const crtb200_kdnode *n = &s->top_nodes[stack[--sp]];
    c->node_tests[c->shadow_query]++;
    if (!box_hit(n, r)) continue;
    if (n->leaf_count) {
      for (uint32_t k = 0; k < n->leaf_count; k++) {
        uint32_t mi = s->top_leaf_refs[n->leaf_start + k];
        hitinfo h = mesh_intersect(c, mi, r, min_t);
        if (h.has) {
          if (!best.has) best = h;
          if (h.t < min_t) { min_t = h.t; best = h; }
        }
      }
    } else {
      if (n->child[0] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[0];
      if (n->child[1] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[1];
    }
  }
  {
    const uint64_t spent = c->node_tests[0] + c->tri_tests[0] - before;
    if (spent > c->max_query_tests) c->max_query_tests = spent;
  }
  if (!best.has) return best;
  const crtb200_material *mat = &s->materials[s->meshes[best.mesh].material];
  int calc_uv = mat->smooth_shading != 0;
  if (mat->texture != CRTB200_INVALID) calc_uv = 1; /* USE_TEXTURES flavour: KDTree.cpp:172-174 */
  best.u = 0;
  best.v = 0;
  if (calc_uv) {
    barycentric(s, best.tri, best.p, &best.u, &best.v);
    if (mat->smooth_shading) { /* KDTree.cpp:180-185 */
      const uint32_t *iv = s->triangle_vertex + 3 * (size_t)best.tri;
      v3 n0 = ld3(s->vertex_normal + 3 * (size_t)iv[0]);
      v3 n1 = ld3(s->vertex_normal + 3 * (size_t)iv[1]);
      v3 n2 = ld3(s->vertex_normal + 3 * (size_t)iv[2]);
      v3 nn = vadd(vadd(vscale(n1, best.u), vscale(n2, best.v)), vscale(n0, (1 - best.u - best.v)));
      best.n = vnorm(nn);
    }
  }
  return best;
}

/* ---- ObjectKDTree::checkForIntersection  src/AccelerationStructure.cpp:56-94 (useGI = false) ---- */
static int scene_occluded(octx *c, const ray_t *r, float distance_to_light) {
  const crtb200_scene *s = c->s;
  int found = 0;
  uint32_t stack[STACK_MAX];
  int sp = 0;
  c->rays[CRTB200_RAY_SHADOW]++;
  c->shadow_query = 1;
  c->shadow_limit = distance_to_light * 1.0001f + 1e-4f + 1e-6f * (fabsf(r->o.x) + fabsf(r->o.y) + fabsf(r->o.z));
  if (s->n_top_nodes == 0) return 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const crtb200_kdnode *n = &s->top_nodes[stack[--sp]];
    c->node_tests[c->shadow_query]++;
    if (!box_hit(n, r)) continue;
    if (n->leaf_count) {
      for (uint32_t k = 0; k < n->leaf_count; k++) {
        uint32_t mi = s->top_leaf_refs[n->leaf_start + k];
        if (r->type == CRTB200_RAY_SHADOW && s->materials[s->meshes[mi].material].type == CRTB200_MAT_REFRACTIVE)
          continue;
        hitinfo h = mesh_intersect(c, mi, r, INFINITY);
        if (h.has && vlen(vsub(h.p, r->o)) <= distance_to_light) found = 1; /* no early out in the reference */
      }
    } else {
      if (n->child[0] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[0];
      if (n->child[1] != CRTB200_INVALID && sp < STACK_MAX) stack[sp++] = n->child[1];
    }
  }
  return found;
}

/* ---- Texture::getColor x4  src/Texture.cpp:14-72 ---- */
static v3 texture_color(const crtb200_scene *s, const crtb200_texture *t, uint32_t tri, float b0, float b1, float b2) {
  switch (t->kind) {
    case CRTB200_TEX_ALBEDO:
      return ld3(t->color_a);
    case CRTB200_TEX_EDGES:
      if (b0 < t->scalar || b1 < t->scalar || b2 < t->scalar) return ld3(t->color_b);
      return ld3(t->color_a);
    default: break;
  }
  const uint32_t *iv = s->triangle_vertex + 3 * (size_t)tri;
  v3 uv0 = s->vertex_uv ? ld3(s->vertex_uv + 3 * (size_t)iv[0]) : V(0, 0, 0);
  v3 uv1 = s->vertex_uv ? ld3(s->vertex_uv + 3 * (size_t)iv[1]) : V(0, 0, 0);
  v3 uv2 = s->vertex_uv ? ld3(s->vertex_uv + 3 * (size_t)iv[2]) : V(0, 0, 0);
  v3 uv = vadd(vadd(sscale(b0, uv1), sscale(b1, uv2)), sscale(b2, uv0)); /* Texture.cpp:34-36, 63-65 */
  if (t->kind == CRTB200_TEX_CHECKER) {
    unsigned x = (unsigned)(uv.x / t->scalar);
    unsigned y = (unsigned)(uv.y / t->scalar);
    return (x % 2 == y % 2) ? ld3(t->color_a) : ld3(t->color_b);
  }
  int w = (int)t->width, h = (int)t->height;
  int x = (int)(uv.x * (float)w);
  int y = (int)((1.0f - uv.y) * (float)h);
  x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
  y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
  return ld3(s->texels + 3 * (t->texel_offset + (size_t)y * w + x));
}

static v3 shoot_ray(octx *c, ray_t *r, unsigned depth);

/* ---- RayTracer::calculateDiffusion  src/RayTracer.cpp:300-330,355 (USE_GI = false) ---- */
static v3 shade_diffuse(octx *c, const hitinfo *h) {
  const crtb200_scene *s = c->s;
  const crtb200_material *mat = &s->materials[s->meshes[h->mesh].material];
  const float PI = (float)M_PI; /* const float PI = M_PIf;  RayTracer.cpp:27 */
  v3 final = V(0, 0, 0);
  for (uint32_t li = 0; li < s->n_lights; li++) {
    const crtb200_light *l = &s->lights[li];
    v3 ld = vsub(ld3(l->position), h->p);
    float dist = vlen(ld);
    float radius = vlen(ld);
    float area = 4 * radius * radius * PI;
    ld = vnorm(ld);
    float angle = stdmax(0.0f, vdot(ld, h->n));
    ray_t sr;
    sr.o = vadd(h->p, vscale(h->n, c->shadow_bias));
    sr.d = ld;
    sr.type = CRTB200_RAY_SHADOW;
    if (!scene_occluded(c, &sr, dist)) {
      float direct = ((float)l->intensity / area * angle);
      v3 base = (mat->texture != CRTB200_INVALID)
                    ? texture_color(s, &s->textures[mat->texture], h->tri, h->u, h->v, 1.0f - h->u - h->v)
                    : ld3(mat->albedo);
      final = vadd(final, sscale(direct, base));
    }
  }
  return final;
}

/* ---- RayTracer::calculateReflection  src/RayTracer.cpp:358-374 ---- */
static v3 shade_reflect(octx *c, const ray_t *r, unsigned depth, const hitinfo *h) {
  const crtb200_material *mat = &c->s->materials[c->s->meshes[h->mesh].material];
  ray_t rr;
  rr.o = vadd(h->p, vscale(h->n, c->reflection_bias));
  rr.d = vnorm(vreflect(r->d, h->n));
  rr.type = CRTB200_RAY_REFLECTION;
  v3 col = shoot_ray(c, &rr, depth + 1);
  return vadd(V(0, 0, 0), V(mat->albedo[0] * col.x, mat->albedo[1] * col.y, mat->albedo[2] * col.z));
}

/* ---- RayTracer::calculateRefraction  src/RayTracer.cpp:375-417 ---- */
static v3 shade_refract(octx *c, const ray_t *r, unsigned depth, const hitinfo *h) {
  const crtb200_material *mat = &c->s->materials[c->s->meshes[h->mesh].material];
  float eta1 = 1.0f, eta2 = mat->ior;
  v3 normal = h->n;
  float idn = vdot(r->d, normal);
  if (idn > 0) {
    float t = eta1; eta1 = eta2; eta2 = t;
    normal = sscale(-1, normal);
    idn = -idn;
  }
  float cos_a = -idn;
  float sin_a = sqrtf(stdmax(0.0f, 1 - cos_a * cos_a));
  ray_t rr;
  rr.o = vadd(h->p, vscale(normal, c->reflection_bias));
  rr.d = vnorm(vreflect(r->d, normal));
  rr.type = CRTB200_RAY_REFLECTION;
  v3 refl = shoot_ray(c, &rr, depth + 1);
  float eta = eta1 / eta2;
  float sin_b = eta * sin_a;
  if (sin_b < 1.0f) {
    float r0 = powf((eta1 - eta2) / (eta1 + eta2), 2);
    float fresnel = r0 + (1 - r0) * powf(1.0f - cos_a, 5);
    float cos_b = sqrtf(stdmax(0.0f, 1 - sin_b * sin_b));
    v3 dir = vsub(sscale(eta, vadd(r->d, sscale(cos_a, normal))), sscale(cos_b, normal));
    ray_t tr;
    tr.o = vsub(h->p, vscale(normal, c->refraction_bias));
    tr.d = vnorm(dir);
    tr.type = CRTB200_RAY_REFRACTION;
    v3 refr = shoot_ray(c, &tr, depth + 1);
    return vadd(sscale(fresnel, refl), sscale(1 - fresnel, refr));
  }
  return refl;
}

/* ---- RayTracer::shootRay  src/RayTracer.cpp:419-451 (boundingType = Tree) ---- */
static v3 shoot_ray(octx *c, ray_t *r, unsigned depth) {
  v3 bg = ld3(c->s->background);
  r->d = vnorm(r->d);
  if (depth > c->max_depth) return bg;
  hitinfo h = scene_intersect(c, r);
  if (!h.has) return bg;
  switch (c->s->materials[c->s->meshes[h.mesh].material].type) {
    case CRTB200_MAT_DIFFUSE: return shade_diffuse(c, &h);
    case CRTB200_MAT_REFLECTIVE: return shade_reflect(c, r, depth, &h);
    case CRTB200_MAT_REFRACTIVE: return shade_refract(c, r, depth, &h);
    default: return bg;
  }
}

/* ---- RayTracer::getRay  src/RayTracer.cpp:61-80 (random = false) ---- */
static ray_t get_ray(const crtb200_scene *s, const crtb200_camera *cam, unsigned row, unsigned col) {
  float x = (float)col + 0.5f;
  float y = (float)row + 0.5f;
  x = x / (float)s->width;
  y = y / (float)s->height;
  x = (2.0f * x) - 1.0f;
  y = 1.0f - (2.0f * y);
  x = x * ((float)s->width / (float)s->height);
  v3 d = V(x, y, -1.0);
  const float *m = cam->rotation; /* Vector * Matrix<3>  include/tracer/Matrix.h:137-142 */
  v3 dr = V(d.x * m[0] + d.y * m[3] + d.z * m[6], d.x * m[1] + d.y * m[4] + d.z * m[7],
            d.x * m[2] + d.y * m[5] + d.z * m[8]);
  ray_t r;
  r.o = ld3(cam->position);
  r.d = vnorm(dr);
  r.type = CRTB200_RAY_PRIMARY;
  return r;
}

static int render_impl(const crtb200_scene *s, const crtb200_camera *cam, const crtb200_options *opt, float *rgb,
                       crtb200_hit *hits, crt_oracle_stats *stats, int threads, const float *skip_mu) {
  if (!s || !cam || !opt || !rgb) return -1;
  crtb200_rect full = {0, 0, s->width, s->height};
  const crtb200_rect *rects = opt->n_rects ? opt->rects : &full;
  uint32_t n_rects = opt->n_rects ? opt->n_rects : 1;
  uint64_t tot_rays[4] = {0, 0, 0, 0}, tot_nodes[2] = {0, 0}, tot_tris[2] = {0, 0}, max_q = 0;
  if (threads <= 0) threads = 1;
  /* ---- RayTracer::renderRectangle  src/RayTracer.cpp:82-112, one work item per image row of a rectangle ---- */
  for (uint32_t ri = 0; ri < n_rects; ri++) {
    const crtb200_rect rc = rects[ri];
    uint32_t row_limit = rc.row + rc.height < s->height ? rc.row + rc.height : s->height;
    uint32_t col_limit = rc.col + rc.width < s->width ? rc.col + rc.width : s->width;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : tot_nodes[:2], tot_tris[:2], tot_rays[:4]) reduction(max : max_q)
    for (uint32_t row = rc.row; row < row_limit; row++) {
      octx c;
      memset(&c, 0, sizeof(c));
      c.s = s;
      c.max_depth = opt->max_depth;
      c.shadow_bias = opt->shadow_bias;
      c.reflection_bias = opt->reflection_bias;
      c.refraction_bias = opt->refraction_bias;
      c.skip_mu = skip_mu;
      for (uint32_t col = rc.col; col < col_limit; col++) {
        ray_t r = get_ray(s, cam, row, col);
        if (hits) {
          ray_t pr = r;
          pr.d = vnorm(pr.d); /* RayTracer.cpp:420 */
          octx tmp = c;
          hitinfo h = scene_intersect(&tmp, &pr);
          crtb200_hit *ho = &hits[(size_t)row * s->width + col];
          if (h.has) {
            ho->mesh = (int32_t)h.mesh;
            ho->triangle = (int32_t)(h.tri - s->meshes[h.mesh].first_triangle);
            ho->t = h.t;
          } else {
            ho->mesh = -1; ho->triangle = -1; ho->t = 0.0f;
          }
        }
        v3 col_out = shoot_ray(&c, &r, 0);
        float *px = rgb + ((size_t)row * s->width + col) * 3;
        px[0] = col_out.x; px[1] = col_out.y; px[2] = col_out.z;
      }
      for (int k = 0; k < 2; k++) {
        tot_nodes[k] += c.node_tests[k];
        tot_tris[k] += c.tri_tests[k];
      }
      for (int k = 0; k < 4; k++) tot_rays[k] += c.rays[k];
      if (c.max_query_tests > max_q) max_q = c.max_query_tests;
    }
  }
  if (stats) {
    stats->rays_primary += tot_rays[CRTB200_RAY_PRIMARY];
    stats->rays_shadow += tot_rays[CRTB200_RAY_SHADOW];
    stats->rays_reflection += tot_rays[CRTB200_RAY_REFLECTION];
    stats->rays_refraction += tot_rays[CRTB200_RAY_REFRACTION];
    stats->node_tests_closest += tot_nodes[0];
    stats->triangle_tests_closest += tot_tris[0];
    stats->node_tests_shadow += tot_nodes[1];
    stats->triangle_tests_shadow += tot_tris[1];
    if (max_q > stats->max_query_tests) stats->max_query_tests = max_q;
  }
  return 0;
}

int crt_oracle_render(const crtb200_scene *s, const crtb200_camera *cam, const crtb200_options *opt, float *rgb,
                      crtb200_hit *hits, crt_oracle_stats *stats, int threads) {
  return render_impl(s, cam, opt, rgb, hits, stats, threads, NULL);
}

/* PPMColor  src/Color.cpp:12-16 */
void crt_oracle_quantize(const float *rgb, size_t n_values, uint8_t *out) {
  for (size_t i = 0; i < n_values; i++) {
    float c = rgb[i];
    c = (c < 0.0f) ? 0.0f : ((1.0f < c) ? 1.0f : c);
    float sc = c * 255;
    unsigned short v = (sc == sc) ? (unsigned short)sc : 0; /* NaN -> cvttss2si 0x80000000 -> low 16 bits 0 */
    out[i] = (uint8_t)v;
  }
}

int crt_oracle_generate_rays(const crtb200_scene *s, const crtb200_camera *cam, float *rays_out) {
  if (!s || !cam || !rays_out) return -1;
  for (uint32_t row = 0; row < s->height; row++)
    for (uint32_t col = 0; col < s->width; col++) {
      ray_t r = get_ray(s, cam, row, col);
      r.d = vnorm(r.d); /* RayTracer.cpp:420 */
      float *o = rays_out + ((size_t)row * s->width + col) * 6;
      o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
    }
  return 0;
}

int crt_oracle_trace_rays(const crtb200_scene *s, const float *rays, uint32_t n, uint32_t ray_type,
                          const float *max_distance, crtb200_hit *hits_out, uint8_t *occluded_out) {
  if (!s || !rays) return -1;
  octx c;
  memset(&c, 0, sizeof(c));
  c.s = s;
  for (uint32_t i = 0; i < n; i++) {
    ray_t r;
    r.o = ld3(rays + 6 * (size_t)i);
    r.d = ld3(rays + 6 * (size_t)i + 3);
    r.type = (int)ray_type;
    if (ray_type == CRTB200_RAY_SHADOW) {
      if (!max_distance || !occluded_out) return -1;
      occluded_out[i] = (uint8_t)scene_occluded(&c, &r, max_distance[i]);
    } else {
      if (!hits_out) return -1;
      hitinfo h = scene_intersect(&c, &r);
      if (h.has) {
        hits_out[i].mesh = (int32_t)h.mesh;
        hits_out[i].triangle = (int32_t)(h.tri - s->meshes[h.mesh].first_triangle);
        hits_out[i].t = h.t;
      } else {
        hits_out[i].mesh = -1; hits_out[i].triangle = -1; hits_out[i].t = 0.0f;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Skip-rule model.  NOT part of the reference restatement: a CPU model of the conservative culling the product's
 * traversal kernels apply by default (csrc/crt_device.cuh node_test), layered on the restated
 * reference walk above so that tests can check on the CPU -- at any size, without GPU time -- that skipping the
 * subtrees the product skips leaves every hit id, t and pixel of the reference walk unchanged.
 *   mu[mesh]  margin in space units such that (i) every leaf L listing a triangle T has bbox(T) inside L inflated by
 *             mu and (ii) every point the triangle test can accept for T is within mu of bbox(T); +inf = never skip.
 *   skip      inside a mesh tree only, for rays with finite o, d, 1/d and no axis-parallel component:
 *             behind: some far slab plane lies more than mu behind the origin        (shadow rays: always; closest
 *                     hit: only once a finite candidate exists)
 *             beyond: some near slab plane lies more than mu beyond the limit        (best finite t / the light)
 * ------------------------------------------------------------------------------------------------------------------ */
static int skip_node(const octx *c, const crtb200_kdnode *n, const ray_t *r, float mesh_mu, float limit, int allow_behind) {
  (void)c;
  const float o[3] = {r->o.x, r->o.y, r->o.z}, d[3] = {r->d.x, r->d.y, r->d.z};
  float inv[3];
  for (int i = 0; i < 3; i++) {
    if (fabsf(d[i]) < FLT_EPSILON) return 0;
    inv[i] = 1.0f / d[i];
    if (!isfinite(o[i]) || !isfinite(d[i]) || !isfinite(inv[i])) return 0;
  }
  const float mu = mesh_mu + (64.0f * FLT_EPSILON) * (fabsf(o[0]) + fabsf(o[1]) + fabsf(o[2]));
  for (int i = 0; i < 3; i++) {
    const float a = (n->box_min[i] - o[i]) * inv[i], b = (n->box_max[i] - o[i]) * inv[i];
    const float tn = fminf(a, b), tf = fmaxf(a, b);
    const float m = mu * fabsf(inv[i]);
    if (allow_behind && tf < -m) return 1;
    if (tn - m > limit) return 1;
  }
  return 0;
}

int crt_oracle_skip_margins(const crtb200_scene *s, float *mu_out) {
  if (!s || !mu_out) return -1;
  for (uint32_t mi = 0; mi < s->n_meshes; mi++) {
    const crtb200_mesh *m = &s->meshes[mi];
    const crtb200_kdnode *nodes = s->mesh_nodes + m->first_node;
    const uint32_t *refs = s->mesh_leaf_refs + m->first_leaf_ref;
    double overhang = 0, slop = 0, absmax = 0;
    int ok = 1;
    for (uint32_t k = 0; k < m->n_nodes && ok; k++) {
      const crtb200_kdnode *n = &nodes[k];
      for (int i = 0; i < 3; i++) {
        if (!(n->box_min[i] <= n->box_max[i])) ok = 0;
        absmax = fmax(absmax, fmax(fabs(n->box_min[i]), fabs(n->box_max[i])));
      }
      if (n->leaf_count == 0) { /* children must lie inside the parent (the skip of a subtree relies on it) */
        for (int side = 0; side < 2; side++) {
          if (n->child[side] == CRTB200_INVALID) continue;
          const crtb200_kdnode *ch = &nodes[n->child[side]];
          for (int i = 0; i < 3; i++)
            if (!(ch->box_min[i] >= n->box_min[i]) || !(ch->box_max[i] <= n->box_max[i])) ok = 0;
        }
        continue;
      }
      for (uint32_t q = 0; q < n->leaf_count; q++) {
        const uint32_t *iv = s->triangle_vertex + 3 * (size_t)(m->first_triangle + refs[n->leaf_start + q]);
        for (int i = 0; i < 3; i++) {
          double lo = INFINITY, hi = -INFINITY;
          for (int v = 0; v < 3; v++) {
            const double x = s->vertex_position[3 * (size_t)iv[v] + i];
            lo = fmin(lo, x);
            hi = fmax(hi, x);
          }
          overhang = fmax(overhang, fmax(n->box_min[i] - lo, hi - n->box_max[i]));
        }
      }
    }
    for (uint32_t t = 0; t < m->n_triangles && ok; t++) {
      const uint32_t *iv = s->triangle_vertex + 3 * (size_t)(m->first_triangle + t);
      double p[3][3];
      for (int v = 0; v < 3; v++)
        for (int i = 0; i < 3; i++) {
          p[v][i] = s->vertex_position[3 * (size_t)iv[v] + i];
          absmax = fmax(absmax, fabs(p[v][i]));
        }
      double e[3][3], len[3];
      for (int k = 0; k < 3; k++) {
        for (int i = 0; i < 3; i++) e[k][i] = p[(k + 1) % 3][i] - p[k][i];
        len[k] = sqrt(e[k][0] * e[k][0] + e[k][1] * e[k][1] + e[k][2] * e[k][2]);
      }
      const double u[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
      const double w[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
      const double cr[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
      const double a2 = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
      const float *nn = s->triangle_normal + 3 * (size_t)(m->first_triangle + t);
      if (!(a2 > 0) || !isfinite(a2)) { /* no plane: only a zero / non-finite normal is harmless (t is never finite) */
        for (int i = 0; i < 3; i++)
          if (isfinite(nn[i]) && nn[i] != 0.0f) ok = 0;
        continue;
      }
      for (int i = 0; i < 3; i++)
        if (!(fabs(nn[i] - cr[i] / a2) <= 1e-3)) ok = 0;
      const double emax = fmax(len[0], fmax(len[1], len[2]));
      slop = fmax(slop, FLT_EPSILON * (1.0 + 8.0 * emax * emax) * (len[0] + len[1] + len[2]) / a2);
    }
    const double mu = overhang + 4.0 * slop + 64.0 * FLT_EPSILON * absmax;
    mu_out[mi] = (ok && isfinite(mu) && mu < 1e30) ? (float)(mu * (1.0 + 1e-6)) : INFINITY;
  }
  return 0;
}

int crt_oracle_render_skip_model(const crtb200_scene *s, const crtb200_camera *cam, const crtb200_options *opt,
                                 const float *mu, float *rgb, crtb200_hit *hits, crt_oracle_stats *stats, int threads) {
  if (!mu) return -1;
  return render_impl(s, cam, opt, rgb, hits, stats, threads, mu);
}
