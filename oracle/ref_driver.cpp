// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Driver around the UNMODIFIED reference renderer (sources stay under /root/reference and are
// compiled from there by oracle/build_ref.sh into oracle/_ref/).  It calls the reference's own
// public API exactly like SourceCode/app/main.cpp:8-20 does:
//   SceneParser::parseScene(file, folder)  ->  RayTracer tracer(scene)  ->  tracer.render(ppm, options)
// and additionally dumps
//   <prefix>.rgbf32 : H*W*3 float32, the colour buffer render() returns (RayTracer.cpp:297)
//   <prefix>.hits   : H*W records {int32 mesh, int32 tri, float t}; primary-ray closest hit obtained with the
//                     reference's own getRay (RayTracer.cpp:61-80), the second normalisation of
//                     shootRay (RayTracer.cpp:420) and accelerationStructure.intersect (KDTree.cpp:127-192)
//   last stdout line: JSON with MEASURE_TIME-equivalent render seconds and the exact ray counts,
//                     counted by ld --wrap interposers on the two tree entry points (trace / shadow).
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "tracer/RayTracer.h"
#undef private
#undef protected
#include "tracer/SceneParser.h"

static std::atomic<unsigned long long> g_closest[5];
static std::atomic<unsigned long long> g_shadow;

using TopTree = KDTree<ObjectKDTreeSubTree>;
extern "C" {
// _ZNK6KDTreeI19ObjectKDTreeSubTreeE9intersectERK3Ray
std::optional<IntersectionInformation> __real__ZNK6KDTreeI19ObjectKDTreeSubTreeE9intersectERK3Ray(const TopTree *self,
                                                                                                 const Ray &ray);
std::optional<IntersectionInformation> __wrap__ZNK6KDTreeI19ObjectKDTreeSubTreeE9intersectERK3Ray(const TopTree *self,
                                                                                                 const Ray &ray) {
  g_closest[static_cast<int>(ray.rayType)].fetch_add(1, std::memory_order_relaxed);
  return __real__ZNK6KDTreeI19ObjectKDTreeSubTreeE9intersectERK3Ray(self, ray);
}
// _ZNK12ObjectKDTree20checkForIntersectionERK3Rayfb
bool __real__ZNK12ObjectKDTree20checkForIntersectionERK3Rayfb(const ObjectKDTree *self, const Ray &ray, float d,
                                                              bool gi);
bool __wrap__ZNK12ObjectKDTree20checkForIntersectionERK3Rayfb(const ObjectKDTree *self, const Ray &ray, float d,
                                                              bool gi) {
  g_shadow.fetch_add(1, std::memory_order_relaxed);
  return __real__ZNK12ObjectKDTree20checkForIntersectionERK3Rayfb(self, ray, d, gi);
}
}

static void usage() {
  std::fprintf(stderr,
               "usage: crt_ref <scene.crtscene> <folder> <out-prefix|-> [--depth N] [--mode M] [--no-hits] "
               "[--no-ppm] [--repeat K] [--cam px py pz r0..r8]\n");
}

int main(int argc, char **argv) {
  if (argc < 4) {
    usage();
    return 2;
  }
  std::string sceneFile = argv[1], folder = argv[2], prefix = argv[3];
  unsigned depth = 5;
  int mode = BVHBucketsThreadPool;
  bool dumpHits = true, writePpm = true, haveCam = false;
  std::string treePath;
  int repeat = 1;
  float cam[12];
  for (int i = 4; i < argc; i++) {
    if (!std::strcmp(argv[i], "--depth") && i + 1 < argc) depth = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--mode") && i + 1 < argc) mode = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--no-hits")) dumpHits = false;
    else if (!std::strcmp(argv[i], "--no-ppm")) writePpm = false;
    else if (!std::strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--dump-tree") && i + 1 < argc) treePath = argv[++i];
    else if (!std::strcmp(argv[i], "--cam") && i + 12 < argc) {
      for (int k = 0; k < 12; k++) cam[k] = std::strtof(argv[++i], nullptr);
      haveCam = true;
    } else {
      usage();
      return 2;
    }
  }
  auto t0 = std::chrono::high_resolution_clock::now();
  Scene scene = SceneParser::parseScene(sceneFile, folder);
  auto t1 = std::chrono::high_resolution_clock::now();
  RayTracer tracer(scene);
  auto t2 = std::chrono::high_resolution_clock::now();
  if (haveCam) {
    tracer.setCamera().setPosition() = Vector(cam[0], cam[1], cam[2]);
    tracer.setCamera().setRotationMatrix() =
        Matrix<3>(std::vector<float>{cam[3], cam[4], cam[5], cam[6], cam[7], cam[8], cam[9], cam[10], cam[11]});
  }
  if (!treePath.empty()) {
    // The reference's own KD trees (KDTree::nodes is protected, KDTree.h:29), node for node, for the build-parity test:
    // per tree: u32 n_nodes, then per node {6 x f32 box, u32 child0, u32 child1, u32 n_indexes, n_indexes x u32}.
    std::ofstream f(treePath, std::ios::binary);
    auto dump = [&f](const auto &nodes) {
      unsigned n = static_cast<unsigned>(nodes.size());
      f.write(reinterpret_cast<const char *>(&n), 4);
      for (const auto &node : nodes) {
        float box[6] = {node.box.minPoint[0], node.box.minPoint[1], node.box.minPoint[2],
                        node.box.maxPoint[0], node.box.maxPoint[1], node.box.maxPoint[2]};
        f.write(reinterpret_cast<const char *>(box), sizeof(box));
        unsigned c[3] = {node.children[0], node.children[1], static_cast<unsigned>(node.indexes.size())};
        f.write(reinterpret_cast<const char *>(c), sizeof(c));
        for (size_t idx : node.indexes) {
          unsigned v = static_cast<unsigned>(idx);
          f.write(reinterpret_cast<const char *>(&v), 4);
        }
      }
    };
    unsigned nMesh = static_cast<unsigned>(tracer.accelerationStructure.container->size());
    f.write(reinterpret_cast<const char *>(&nMesh), 4);
    for (const auto &sub : *tracer.accelerationStructure.container) dump(sub.tree.nodes);
    dump(tracer.accelerationStructure.nodes);
  }
  const unsigned W = tracer.scene.sceneSettings.image.width, H = tracer.scene.sceneSettings.image.height;
  RenderOptions options{static_cast<RenderOptimization>(mode), depth, false};

  // The reference prints its MEASURE_TIME figure ("<seconds>s\n", RayTracer.cpp:289-293) on std::cout while the
  // progress bar goes through printf; capture std::cout to read exactly the number the reference reports.
  double best = 1e300;
  std::vector<double> allSeconds;
  std::vector<std::vector<Color>> buffer;
  unsigned long long counts[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < repeat; r++) {
    for (auto &c : g_closest) c = 0;
    g_shadow = 0;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    auto s = std::chrono::high_resolution_clock::now();
    buffer = tracer.render((writePpm && prefix != "-" && r == 0) ? prefix + ".ppm" : std::string(""), options);
    auto e = std::chrono::high_resolution_clock::now();
    std::cout.rdbuf(old);
    double seconds = std::chrono::duration<double>(e - s).count();
    {
      // last token of the form "<float>s"
      std::string text = captured.str();
      size_t pos = text.rfind("s\n");
      if (pos != std::string::npos) {
        size_t b = text.rfind('\n', pos);
        b = (b == std::string::npos) ? 0 : b + 1;
        double v = std::strtod(text.substr(b, pos - b).c_str(), nullptr);
        if (v > 0) seconds = v;
      }
    }
    for (int k = 0; k < 5; k++) counts[k] = g_closest[k];
    counts[5] = g_shadow;
    best = std::min(best, seconds);
    allSeconds.push_back(seconds);
  }
  std::printf("\n");

  if (prefix != "-") {
    std::ofstream f(prefix + ".rgbf32", std::ios::binary);
    for (unsigned row = 0; row < H; row++)
      for (unsigned col = 0; col < W; col++) {
        float px[3] = {buffer[row][col][0], buffer[row][col][1], buffer[row][col][2]};
        f.write(reinterpret_cast<const char *>(px), sizeof(px));
      }
  }
  if (dumpHits && prefix != "-") {
    struct Rec {
      int mesh, tri;
      float t;
    };
    std::vector<Rec> recs(static_cast<size_t>(W) * H);
    const Mesh *mesh0 = &tracer.scene.objects[0];
    unsigned nThreads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (unsigned tid = 0; tid < nThreads; tid++) {
      pool.emplace_back([&, tid]() {
        for (unsigned row = tid; row < H; row += nThreads)
          for (unsigned col = 0; col < W; col++) {
            Ray ray = tracer.getRay(row, col, false);
            ray.direction.normalize();  // RayTracer.cpp:420
            auto info = tracer.accelerationStructure.intersect(ray);
            Rec rec{-1, -1, 0.0f};
            if (info.has_value()) {
              rec.mesh = static_cast<int>(info->object - mesh0);
              rec.tri = static_cast<int>(info->triangle - &info->object->triangles[0]);
              rec.t = info->intersection.distance;
            }
            recs[static_cast<size_t>(row) * W + col] = rec;
          }
      });
    }
    for (auto &t : pool) t.join();
    std::ofstream f(prefix + ".hits", std::ios::binary);
    f.write(reinterpret_cast<const char *>(recs.data()), recs.size() * sizeof(Rec));
  }
  size_t tris = 0;
  for (auto &o : tracer.scene.objects) tris += o.triangles.size();
  std::string all = "[";
  for (size_t i = 0; i < allSeconds.size(); i++) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%s%.6f", i ? ", " : "", allSeconds[i]);
    all += buf;
  }
  all += "]";
  std::printf(
      "{\"width\": %u, \"height\": %u, \"triangles\": %zu, \"meshes\": %zu, \"threads\": %u, \"parse_s\": %.6f, "
      "\"build_s\": %.6f, \"render_s\": %.6f, \"render_all_s\": %s, \"rays\": {\"primary\": %llu, \"shadow\": %llu, \"reflection\": %llu, "
      "\"refraction\": %llu}}\n",
      W, H, tris, tracer.scene.objects.size(), std::thread::hardware_concurrency(),
      std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), best, all.c_str(),
      counts[PrimaryRay], counts[5], counts[ReflectionRay], counts[RefractionRay]);
  return 0;
}
