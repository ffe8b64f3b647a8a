#include "crt_scene.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <thread>

#include "crt_json.hpp"

namespace crt {

// ---- Camera (Camera.cpp:10-70).  degrees -> radians uses pi ~ 22/7 like the reference. ----
static float degreesToRadians(float degrees) { return degrees * (22 / (7 * 180.0f)); }

Camera &Camera::truck(const Vector &direction) {
  position += direction * rotationMatrix;
  return *this;
}
Camera &Camera::pan(float degrees) {
  const float r = degreesToRadians(degrees);
  Matrix3 rot{cosf(r), 0.0f, -sinf(r), 0.0f, 1.0f, 0.0f, sinf(r), 0.0f, cosf(r)};
  rotationMatrix *= rot;
  return *this;
}
Camera &Camera::roll(float degrees) {
  const float r = degreesToRadians(degrees);
  Matrix3 rot{cosf(r), -sinf(r), 0.0f, sinf(r), cosf(r), 0.0f, 0.0f, 0.0f, 1.0f};
  rotationMatrix *= rot;
  return *this;
}
Camera &Camera::tilt(float degrees) {
  const float r = degreesToRadians(degrees);
  Matrix3 rot{1.0f, 0.0f, 0.0f, 0.0f, cosf(r), -sinf(r), 0.0f, sinf(r), cosf(r)};
  rotationMatrix *= rot;
  return *this;
}

// ---- Mesh ----
Mesh::Mesh(unsigned materialIndex_, std::vector<Vector> positions_, std::vector<uint32_t> indices_,
           std::vector<Vector> uvs_)
    : materialIndex(materialIndex_), positions(std::move(positions_)), uvs(std::move(uvs_)), indices(std::move(indices_)) {
  normals.assign(positions.size(), Vector());
  const size_t nt = indices.size() / 3;
  faceNormals.resize(nt);
  for (size_t t = 0; t < nt; t++) {
    const uint32_t i0 = indices[3 * t], i1 = indices[3 * t + 1], i2 = indices[3 * t + 2];
    if (i0 >= positions.size() || i1 >= positions.size() || i2 >= positions.size())
      throw std::runtime_error("mesh triangle index out of range");
    Vector e1 = positions[i1] - positions[i0];
    Vector e2 = positions[i2] - positions[i0];
    Vector n = e1 * e2;  // cross
    n.normalize();
    faceNormals[t] = n;
    normals[i0] += n;
    normals[i1] += n;
    normals[i2] += n;
  }
  for (auto &n : normals) n.normalize();
}

// ---- image decoding for bitmap textures (the reference uses vendored stb_image, Texture.cpp:46-60) ----
static unsigned be32(const unsigned char *p) { return (unsigned(p[0]) << 24) | (unsigned(p[1]) << 16) | (unsigned(p[2]) << 8) | p[3]; }

static bool loadPng(const std::vector<unsigned char> &file, int &w, int &h, int &ch, std::vector<unsigned char> &out,
                    std::string &err) {
  size_t pos = 8;
  int bitDepth = 0, colorType = 0, interlace = 0;
  bool haveHeader = false;
  w = h = 0;
  std::vector<unsigned char> idat, palette;
  while (pos + 12 <= file.size()) {
    unsigned len = be32(&file[pos]);
    const unsigned char *tag = &file[pos + 4];
    const unsigned char *data = &file[pos + 8];
    if (pos + 12 + len > file.size()) break;
    if (!std::memcmp(tag, "IHDR", 4)) {
      if (len < 13) {
        err = "PNG: truncated IHDR";
        return false;
      }
      const unsigned uw = be32(data), uh = be32(data + 4);
      if (uw == 0 || uh == 0 || uw > 32768u || uh > 32768u) {
        err = "PNG: image size out of range (1..32768)";
        return false;
      }
      haveHeader = true;
      w = int(uw);
      h = int(uh);
      bitDepth = data[8];
      colorType = data[9];
      interlace = data[12];
    } else if (!std::memcmp(tag, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!std::memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(tag, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (!haveHeader) {
    err = "PNG: no IHDR chunk";
    return false;
  }
  if (bitDepth != 8 || interlace != 0) {
    err = "PNG: only 8-bit non-interlaced images are supported";
    return false;
  }
  int srcCh = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : colorType == 6 ? 4 : 0;
  if (!srcCh) {
    err = "PNG: bad colour type";
    return false;
  }
  const size_t stride = size_t(w) * srcCh;
  std::vector<unsigned char> raw((stride + 1) * size_t(h));
  uLongf rawLen = raw.size();
  if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size()) {
    err = "PNG: inflate failed";
    return false;
  }
  std::vector<unsigned char> img(stride * size_t(h));
  for (int y = 0; y < h; y++) {
    const unsigned char *src = &raw[(stride + 1) * size_t(y)];
    unsigned char *dst = &img[stride * size_t(y)];
    const unsigned char *up = y ? dst - stride : nullptr;
    const int f = src[0];
    for (size_t i = 0; i < stride; i++) {
      int a = i >= size_t(srcCh) ? dst[i - srcCh] : 0;
      int b = up ? up[i] : 0;
      int c = (up && i >= size_t(srcCh)) ? up[i - srcCh] : 0;
      int x = src[i + 1];
      switch (f) {
        case 0: break;
        case 1: x += a; break;
        case 2: x += b; break;
        case 3: x += (a + b) / 2; break;
        case 4: {
          int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: err = "PNG: bad filter"; return false;
      }
      dst[i] = static_cast<unsigned char>(x);
    }
  }
  if (colorType == 3) {  // palette -> RGB, like stb_image with req_comp = 0
    ch = 3;
    out.resize(size_t(w) * h * 3);
    for (size_t i = 0; i < size_t(w) * h; i++)
      for (int k = 0; k < 3; k++) out[3 * i + k] = (size_t(img[i]) * 3 + k < palette.size()) ? palette[img[i] * 3 + k] : 0;
  } else {
    ch = srcCh;
    out.swap(img);
  }
  return true;
}

bool loadImage8(const std::string &path, int &w, int &h, int &ch, std::vector<unsigned char> &out, std::string &err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    err = "cannot open " + path;
    return false;
  }
  std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (file.size() > 8 && !std::memcmp(file.data(), sig, 8)) return loadPng(file, w, h, ch, out, err);
  if (file.size() > 2 && file[0] == 'P' && file[1] == '6') {
    size_t pos = 2;
    int vals[3], got = 0;
    while (got < 3 && pos < file.size()) {
      while (pos < file.size() && std::isspace(file[pos])) pos++;
      if (pos < file.size() && file[pos] == '#') {
        while (pos < file.size() && file[pos] != '\n') pos++;
        continue;
      }
      int v = 0;
      while (pos < file.size() && std::isdigit(file[pos])) v = v * 10 + (file[pos++] - '0');
      vals[got++] = v;
    }
    pos++;
    w = vals[0];
    h = vals[1];
    ch = 3;
    if (vals[2] != 255 || pos + size_t(w) * h * 3 > file.size()) {
      err = "PPM: only maxval 255 supported";
      return false;
    }
    out.assign(file.begin() + pos, file.begin() + pos + size_t(w) * h * 3);
    return true;
  }
  err = "unsupported image format (PNG / P6 only): " + path;
  return false;
}

// ---- SceneParser (SceneParser.cpp:39-322; key names SceneParser.cpp:17-35) ----
using json::Value;

static const Value &member(const Value &o, const char *key) {
  const Value *v = o.find(key);
  if (!v) throw std::runtime_error(std::string("crtscene: missing key \"") + key + "\"");
  return *v;
}
static Vector vec3(const Value &a, size_t off = 0) {
  if (!a.isArray() || a.size() < off + 3) throw std::runtime_error("crtscene: expected 3 numbers");
  // GetFloat() == static_cast<float>(GetDouble())  (SceneParser.cpp:79)
  return Vector(static_cast<float>(a.number(off)), static_cast<float>(a.number(off + 1)), static_cast<float>(a.number(off + 2)));
}

Scene SceneParser::parseSceneText(const char *begin, const char *end, const std::string &sceneFolder) {
  json::Parser parser(begin, end);
  Value doc = parser.parse();
  Scene scene;

  // settings (SceneParser.cpp:88-114)
  {
    const Value &st = member(doc, "settings");
    scene.sceneSettings.sceneBackgroundColor = vec3(member(st, "background_color"));
    const Value &im = member(st, "image_settings");
    unsigned hc = std::thread::hardware_concurrency();
    unsigned bucket = (hc == 1) ? 1 : hc * 6;
    if (const Value *b = im.find("bucket_size"))
      if (b->kind == Value::Number && b->isInt) bucket = static_cast<unsigned>(static_cast<int>(b->num));
    scene.sceneSettings.bucketSize = bucket;
    scene.sceneSettings.image.width = static_cast<unsigned>(member(im, "width").num);
    scene.sceneSettings.image.height = static_cast<unsigned>(member(im, "height").num);
  }
  // camera (SceneParser.cpp:116-130)
  if (const Value *cam = doc.find("camera")) {
    if (cam->kind == Value::Object) {
      scene.camera.position = vec3(member(*cam, "position"));
      const Value &m = member(*cam, "matrix");
      if (!m.isArray() || m.size() < 9) throw std::runtime_error("crtscene: camera.matrix needs 9 numbers");
      for (int k = 0; k < 9; k++) scene.camera.rotationMatrix.m[k / 3][k % 3] = static_cast<float>(m.number(k));
    }
  }
  // textures (SceneParser.cpp:149-209)
  if (const Value *texs = doc.find("textures")) {
    if (texs->kind == Value::Array) {
      for (const Value &t : texs->arr) {
        Texture tex;
        tex.name = member(t, "name").str;
        const std::string &type = member(t, "type").str;
        if (type == "albedo") {
          tex.kind = AlbedoTextureKind;
          tex.colorA = vec3(member(t, "albedo"));
        } else if (type == "edges") {
          tex.kind = EdgeTextureKind;
          tex.colorA = vec3(member(t, "inner_color"));
          tex.colorB = vec3(member(t, "edge_color"));
          tex.scalar = static_cast<float>(member(t, "edge_width").num);
        } else if (type == "checker") {
          tex.kind = CheckerTextureKind;
          tex.colorA = vec3(member(t, "color_A"));
          tex.colorB = vec3(member(t, "color_B"));
          tex.scalar = static_cast<float>(member(t, "square_size").num);
        } else if (type == "bitmap") {
          tex.kind = BitmapTextureKind;
          // folder + file_path with no separator, SceneParser.cpp:201
          std::string path = sceneFolder + member(t, "file_path").str;
          std::vector<unsigned char> px;
          std::string err;
          if (!loadImage8(path, tex.width, tex.height, tex.channels, px, err))
            throw std::runtime_error("crtscene: bitmap texture: " + err);
          if (tex.channels < 3) throw std::runtime_error("crtscene: bitmap texture needs >= 3 channels");
          const size_t n = size_t(tex.width) * tex.height;
          tex.buffer.resize(n);
          const float k = 1.0f / 255.0f;  // Texture.cpp:55-59
          for (size_t i = 0; i < n; i++)
            tex.buffer[i] = Color(static_cast<float>(px[tex.channels * i + 0]) * k,
                                  static_cast<float>(px[tex.channels * i + 1]) * k,
                                  static_cast<float>(px[tex.channels * i + 2]) * k);
        } else {
          throw std::runtime_error("Invalid material");  // SceneParser.cpp:203 (sic)
        }
        scene.textures.push_back(std::move(tex));
      }
    }
  }
  // materials (SceneParser.cpp:211-271)
  if (const Value *mats = doc.find("materials")) {
    if (mats->kind == Value::Array) {
      for (const Value &m : mats->arr) {
        Material mat;
        mat.albedo = Albedo(0, 0, 0);
        mat.ior = 0;  // SceneParser.cpp:222
        const std::string &type = member(m, "type").str;
        if (type == "diffuse")
          mat.type = Diffuse;
        else if (type == "reflective")
          mat.type = Reflective;
        else if (type == "refractive") {
          mat.type = Refractive;
          mat.ior = static_cast<float>(member(m, "ior").num);
        } else if (type == "constant")
          mat.type = Constant;
        else
          throw std::runtime_error("Invalid material");  // SceneParser.cpp:237
        mat.smoothShading = member(m, "smooth_shading").b;
        const Value *alb = m.find("albedo");
        if (alb && alb->kind == Value::String) {
          // USE_TEXTURES flavour: "albedo" names a texture (SceneParser.cpp:241-251)
          for (size_t k = 0; k < scene.textures.size(); k++)
            if (scene.textures[k].name == alb->str) {
              mat.texture = static_cast<int>(k);
              break;
            }
          if (mat.texture < 0) throw std::runtime_error("crtscene: unknown texture \"" + alb->str + "\"");
        } else if (alb && alb->isArray() && mat.type != Refractive) {
          mat.albedo = vec3(*alb);  // SceneParser.cpp:260-264
        }
        scene.materials.push_back(mat);
      }
    }
  }
  // lights (SceneParser.cpp:132-147)
  if (const Value *ls = doc.find("lights")) {
    if (ls->kind == Value::Array)
      for (const Value &l : ls->arr) {
        Light light;
        light.position = vec3(member(l, "position"));
        light.intensity = static_cast<unsigned>(member(l, "intensity").num);
        scene.lights.push_back(light);
      }
  }
  // objects (SceneParser.cpp:273-322)
  if (const Value *objs = doc.find("objects")) {
    if (objs->kind == Value::Array) {
      scene.objects.reserve(objs->arr.size());
      for (const Value &o : objs->arr) {
        unsigned matIndex = static_cast<unsigned>(member(o, "material_index").num);
        if (matIndex >= scene.materials.size()) throw std::runtime_error("crtscene: material_index out of range");
        const Value &vs = member(o, "vertices");
        if (!vs.isArray() || vs.size() % 3) throw std::runtime_error("crtscene: vertices must be 3n numbers");
        std::vector<Vector> pos(vs.size() / 3);
        for (size_t i = 0; i < pos.size(); i++) pos[i] = vec3(vs, 3 * i);
        std::vector<Vector> uvs;
        if (const Value *uv = o.find("uvs")) {
          if (uv->isArray() && uv->size() % 3 == 0) {
            uvs.assign(pos.size(), Vector());
            for (size_t i = 0; i < uv->size() / 3 && i < pos.size(); i++) uvs[i] = vec3(*uv, 3 * i);
          }
        }
        const Value &ts = member(o, "triangles");
        if (!ts.isArray() || ts.size() % 3) throw std::runtime_error("crtscene: triangles must be 3m indices");
        std::vector<uint32_t> idx(ts.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = static_cast<uint32_t>(ts.number(i));
        scene.objects.emplace_back(matIndex, std::move(pos), std::move(idx), std::move(uvs));
      }
    }
  }
  return scene;
}

Scene SceneParser::parseScene(const std::string &pathToScene, const std::string &sceneFolder) {
  const std::string path = (sceneFolder.empty() ? "" : sceneFolder + "/") + pathToScene;  // SceneParser.cpp:40
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open scene file " + path);
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::string text(static_cast<size_t>(n), '\0');
  size_t got = std::fread(text.data(), 1, static_cast<size_t>(n), f);
  std::fclose(f);
  return parseSceneText(text.data(), text.data() + got, sceneFolder);
}

}  // namespace crt
