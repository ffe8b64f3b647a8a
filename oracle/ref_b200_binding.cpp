// TEST INFRASTRUCTURE ONLY -- INTEGRATION.md section B compiled for real.
//
// The binding a maintainer of the reference adds to route RayTracer::render to libcrtb200.so, built here AGAINST THE
// UNMODIFIED REFERENCE SOURCES (compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/crt_ref_b200): the
// reference's own SceneParser loads the scene, the reference's own RayTracer constructor builds its AccelerationStructure
// (KDTree::build), this file flattens those objects into the C ABI of include/crtb200.h, renders through
// crtb200_render, copies the result into the reference's colorBuffer and lets the reference's own exportPPM write it.
// A GPU test compares that PPM byte for byte with the one crt_ref (the reference rendering on the CPU) writes.
//
// Access: RayTracer::scene / accelerationStructure / colorBuffer / camera are private and KDTree::nodes / container
// protected (RayTracer.h:61-69, KDTree.h:29-30).  A maintainer adds `friend` declarations or makes the binding a member;
// this stand-alone build reaches them with the `#define private public` trick of oracle/ref_driver.cpp instead, so the
// reference sources stay untouched.
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <optional>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "tracer/RayTracer.h"
#undef private
#undef protected
#include "tracer/SceneParser.h"

#include "../include/crtb200.h"

namespace {

struct B200State {  // what INTEGRATION.md section B adds to RayTracer as `B200State *b200`
  crtb200_ctx *ctx = nullptr;
  std::vector<float> pos, nrm, triNormal;
  std::vector<uint32_t> triVertex, meshRefs, topRefs;
  std::vector<crtb200_mesh> meshes;
  std::vector<crtb200_material> materials;
  std::vector<crtb200_light> lights;
  std::vector<crtb200_kdnode> meshNodes, topNodes;
  std::vector<float> frame;  // H*W*3, persists like colorBuffer (RayTracer.h:69)
};

// KDTree<T>::nodes -> crtb200_kdnode[], in the reference's numbering (KDTree.h:16-21, 33-38)
template <class Tree>
void flattenTree(const Tree &tree, std::vector<crtb200_kdnode> &nodes, std::vector<uint32_t> &refs) {
  const size_t refBase = refs.size();  // leaf_start is relative to the tree's first reference
  for (const auto &n : tree.nodes) {
    crtb200_kdnode o{};
    for (int i = 0; i < 3; i++) {
      o.box_min[i] = n.box.minPoint[i];
      o.box_max[i] = n.box.maxPoint[i];
    }
    o.child[0] = n.children[0];  // INVALID_INDEX == CRTB200_INVALID
    o.child[1] = n.children[1];
    o.leaf_start = static_cast<uint32_t>(refs.size() - refBase);
    o.leaf_count = static_cast<uint32_t>(n.indexes.size());
    for (size_t idx : n.indexes) refs.push_back(static_cast<uint32_t>(idx));
    nodes.push_back(o);
  }
}

// RayTracer::initB200(): called once after RayTracer::RayTracer (RayTracer.cpp:45-51) has built the trees
void initB200(RayTracer &rt, B200State &s, int nDevices) {
  const Scene &scene = rt.scene;
  uint32_t vBase = 0, tBase = 0;
  for (size_t m = 0; m < scene.objects.size(); m++) {  // Mesh (Scene.h:25-38)
    const Mesh &mesh = scene.objects[m];
    crtb200_mesh fm{};
    fm.material = static_cast<uint32_t>(&mesh.material - &scene.materials[0]);
    fm.first_vertex = vBase;
    fm.n_vertices = static_cast<uint32_t>(mesh.vertices.size());
    fm.first_triangle = tBase;
    fm.n_triangles = static_cast<uint32_t>(mesh.triangles.size());
    for (const Vertex &v : mesh.vertices)
      for (int i = 0; i < 3; i++) {
        s.pos.push_back(v.position[i]);
        s.nrm.push_back(v.normal[i]);
      }
    for (const Triangle &t : mesh.triangles)
      for (int k = 0; k < 3; k++) {
        s.triVertex.push_back(vBase + static_cast<uint32_t>(&t[k] - &mesh.vertices[0]));  // Vertex* -> index
        s.triNormal.push_back(t.getTriangleNormal()[k]);                                   // Triangle.cpp:13-16
      }
    const ObjectKDTreeSubTree &sub = rt.accelerationStructure.container->at(m);  // AccelerationStructure.h:15-20
    fm.first_node = static_cast<uint32_t>(s.meshNodes.size());
    fm.first_leaf_ref = static_cast<uint32_t>(s.meshRefs.size());
    flattenTree(sub.tree, s.meshNodes, s.meshRefs);
    fm.n_nodes = static_cast<uint32_t>(s.meshNodes.size()) - fm.first_node;
    fm.n_leaf_refs = static_cast<uint32_t>(s.meshRefs.size()) - fm.first_leaf_ref;
    s.meshes.push_back(fm);
    vBase += fm.n_vertices;
    tBase += fm.n_triangles;
  }
  flattenTree(rt.accelerationStructure, s.topNodes, s.topRefs);  // ObjectKDTree (AccelerationStructure.h:21-33)
  for (const Material &m : scene.materials) {                    // Material.h:9-30 (non-USE_TEXTURES flavour)
    crtb200_material fm{};
    fm.type = static_cast<uint32_t>(m.type);
    fm.smooth_shading = m.smoothShading ? 1u : 0u;
    fm.texture = CRTB200_INVALID;
    fm.ior = m.ior;
    for (int i = 0; i < 3; i++) fm.albedo[i] = m.albedo[i];
    s.materials.push_back(fm);
  }
  for (const Light &l : scene.lights) {  // Scene.h:20-23
    crtb200_light fl{};
    for (int i = 0; i < 3; i++) fl.position[i] = l.position[i];
    fl.intensity = l.intentsity;
    s.lights.push_back(fl);
  }
  crtb200_scene a{};
  a.abi_version = CRTB200_ABI_VERSION;
  a.width = scene.sceneSettings.image.width;
  a.height = scene.sceneSettings.image.height;
  for (int i = 0; i < 3; i++) a.background[i] = scene.sceneSettings.sceneBackgroundColor[i];
  a.n_vertices = vBase;
  a.vertex_position = s.pos.data();
  a.vertex_normal = s.nrm.data();
  a.vertex_uv = nullptr;
  a.n_triangles = tBase;
  a.triangle_vertex = s.triVertex.data();
  a.triangle_normal = s.triNormal.data();
  a.n_meshes = static_cast<uint32_t>(s.meshes.size());
  a.meshes = s.meshes.data();
  a.n_materials = static_cast<uint32_t>(s.materials.size());
  a.materials = s.materials.data();
  a.n_lights = static_cast<uint32_t>(s.lights.size());
  a.lights = s.lights.data();
  a.n_mesh_nodes = static_cast<uint32_t>(s.meshNodes.size());
  a.mesh_nodes = s.meshNodes.data();
  a.n_mesh_leaf_refs = static_cast<uint32_t>(s.meshRefs.size());
  a.mesh_leaf_refs = s.meshRefs.data();
  a.n_top_nodes = static_cast<uint32_t>(s.topNodes.size());
  a.top_nodes = s.topNodes.data();
  a.n_top_leaf_refs = static_cast<uint32_t>(s.topRefs.size());
  a.top_leaf_refs = s.topRefs.data();
  std::vector<int> devices(static_cast<size_t>(std::max(1, nDevices)));
  for (size_t i = 0; i < devices.size(); i++) devices[i] = static_cast<int>(i);
  if (crtb200_create_multi(devices.data(), static_cast<int>(devices.size()), &s.ctx)) throw std::runtime_error(crtb200_last_error());
  if (crtb200_upload_scene(s.ctx, &a)) throw std::runtime_error(crtb200_last_error_ctx(s.ctx));
  s.frame.assign(size_t(a.width) * a.height * 3, 0.0f);
}

// RayTracer::renderB200(): the new `case B200Wavefront:` of the scheduler switch (RayTracer.cpp:209-286)
void renderB200(RayTracer &rt, B200State &s, unsigned maxDepth) {
  const unsigned W = rt.scene.sceneSettings.image.width, H = rt.scene.sceneSettings.image.height;
  const unsigned short rectangleCount = static_cast<unsigned short>(rt.scene.sceneSettings.bucketSize);  // RayTracer.cpp:274
  std::vector<crtb200_rect> rects;  // the grid renderBucketsThreadpool hands to renderRectangle (RayTracer.cpp:143-152, 84-85)
  const unsigned ny = std::max(1u, static_cast<unsigned>(std::sqrt(rectangleCount))), nx = rectangleCount / ny;
  const unsigned w = W / nx, h = H / ny;
  for (unsigned i = 0; i < rectangleCount; i++) {
    const unsigned row = (i / nx) * h, col = (i * w) % W;
    const unsigned rowLimit = std::min(H, row + h), colLimit = std::min(W, col + w);
    if (row < rowLimit && col < colLimit) rects.push_back({row, col, colLimit - col, rowLimit - row});
  }
  crtb200_camera cam;
  for (int i = 0; i < 3; i++) cam.position[i] = rt.camera.getPosition()[i];
  for (int i = 0; i < 9; i++) cam.rotation[i] = rt.camera.getRotationMatrix()[i / 3][i % 3];
  crtb200_options o{};
  o.max_depth = maxDepth;
  o.shadow_bias = o.reflection_bias = o.refraction_bias = 1e-4f;  // RenderOptions defaults (RayTracer.h:30-33)
  o.n_rects = static_cast<uint32_t>(rects.size());
  o.rects = rects.data();
  if (crtb200_render(s.ctx, &cam, &o, s.frame.data(), nullptr, nullptr, nullptr)) throw std::runtime_error(crtb200_last_error_ctx(s.ctx));
  for (unsigned r = 0; r < H; r++)  // back into colorBuffer: exportPPM and the return value of render() read it
    for (unsigned c = 0; c < W; c++) {
      const float *p = &s.frame[(size_t(r) * W + c) * 3];
      rt.colorBuffer[r][c] = Color(p[0], p[1], p[2]);
    }
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: crt_ref_b200 <scene.crtscene> <folder> <out.ppm> [--depth N] [--devices N] [--cam px py pz r0..r8]\n");
    return 2;
  }
  unsigned depth = 5;
  int nDevices = 1;
  bool haveCam = false;
  float cam[12];
  for (int i = 4; i < argc; i++) {
    if (!std::strcmp(argv[i], "--depth") && i + 1 < argc) depth = static_cast<unsigned>(std::atoi(argv[++i]));
    else if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) nDevices = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--cam") && i + 12 < argc) {
      for (int k = 0; k < 12; k++) cam[k] = std::strtof(argv[++i], nullptr);
      haveCam = true;
    }
  }
  try {
    Scene scene = SceneParser::parseScene(argv[1], argv[2]);  // the reference's loader
    RayTracer tracer(scene);                                  // the reference's AABB + KD-tree build
    if (haveCam) {
      tracer.setCamera().setPosition() = Vector(cam[0], cam[1], cam[2]);
      tracer.setCamera().setRotationMatrix() =
          Matrix<3>(std::vector<float>{cam[3], cam[4], cam[5], cam[6], cam[7], cam[8], cam[9], cam[10], cam[11]});
    }
    B200State state;
    initB200(tracer, state, nDevices);
    renderB200(tracer, state, depth);
    tracer.exportPPM(argv[3], tracer.colorBuffer);  // the reference's own P3 writer (RayTracer.cpp:540-552)
    crtb200_destroy(state.ctx);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "crt_ref_b200: %s\n", e.what());
    return 1;
  } catch (const char *e) {
    std::fprintf(stderr, "crt_ref_b200: %s\n", e);
    return 1;
  }
  return 0;
}
