// Scene model of the host front end: the reference's Scene / Mesh / Material / Texture / Light / Camera /
// SceneParser API surface (include/tracer/Scene.h:8-69, Material.h:7-30, Texture.h:8-59, Camera.h:5-21,
// SceneParser.h:10-44), re-designed index-linked instead of pointer-linked so it flattens
// straight into the C ABI's SoA arrays (include/crtb200.h).  The reference's compile-time USE_TEXTURES flavour
// (CMakeLists.txt:18-19) is a run-time property here: a scene has textures iff its JSON has them.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "crt_math.hpp"

namespace crt {

struct Image {
  unsigned width = 0, height = 0;
};
struct SceneSettings {
  Color sceneBackgroundColor;
  Image image;
  unsigned bucketSize = 1;  // NUMBER of buckets
};
struct Light {
  Vector position;
  unsigned intensity = 0;
};

enum MaterialType { Diffuse, Reflective, Constant, Refractive };  // Material.h:7 (same numeric values)
enum TextureKind { AlbedoTextureKind, EdgeTextureKind, CheckerTextureKind, BitmapTextureKind };

struct Texture {
  std::string name;
  TextureKind kind = AlbedoTextureKind;
  Color colorA, colorB;  // albedo | inner,edge | A,B
  float scalar = 0;      // edge width | square size
  int width = 0, height = 0, channels = 0;
  std::vector<Color> buffer;  // bitmap texels / 255 (Texture.cpp:46-60)
};

struct Material {
  Albedo albedo;
  MaterialType type = Diffuse;
  bool smoothShading = false;
  float ior = 1.0f;
  int texture = -1;  // index into Scene::textures, -1 = none (non-USE_TEXTURES flavour)
};

struct Camera {
  Vector position;
  Matrix3 rotationMatrix = Matrix3::identity();
  Camera() = default;
  explicit Camera(const Vector &p) : position(p) {}
  const Vector &getPosition() const { return position; }
  Vector &setPosition() { return position; }
  const Matrix3 &getRotationMatrix() const { return rotationMatrix; }
  Matrix3 &setRotationMatrix() { return rotationMatrix; }
  Camera &truck(const Vector &direction);
  Camera &pan(float degrees);
  Camera &tilt(float degrees);
  Camera &roll(float degrees);
};

// Mesh::Mesh (Scene.cpp:5-30): face normals, then area-unweighted vertex normals.
struct Mesh {
  unsigned materialIndex = 0;
  std::vector<Vector> positions, normals, uvs;  // uvs empty when the object has none
  std::vector<uint32_t> indices;                // 3 per triangle (mesh-local vertex ids)
  std::vector<Vector> faceNormals;              // Triangle::normal (Triangle.cpp:13-16)
  Mesh() = default;
  Mesh(unsigned materialIndex, std::vector<Vector> positions, std::vector<uint32_t> indices,
       std::vector<Vector> uvs = {});
  size_t triangleCount() const { return indices.size() / 3; }
};

struct Scene {
  SceneSettings sceneSettings;
  Camera camera;
  std::vector<Texture> textures;
  std::vector<Material> materials;
  std::vector<Light> lights;
  std::vector<Mesh> objects;
  size_t triangleCount() const {
    size_t n = 0;
    for (auto &o : objects) n += o.triangleCount();
    return n;
  }
};

class SceneParser {
 public:
  // SceneParser::parseScene(pathToScene, sceneFolder)  -- SceneParser.cpp:39-66
  static Scene parseScene(const std::string &pathToScene, const std::string &sceneFolder = "");
  static Scene parseSceneText(const char *begin, const char *end, const std::string &sceneFolder = "");
};

// decodes 8-bit PNG (gray / RGB / palette / gray+alpha / RGBA, non-interlaced) or binary PPM (P6)
bool loadImage8(const std::string &path, int &width, int &height, int &channels, std::vector<unsigned char> &pixels,
                std::string &error);

}  // namespace crt
