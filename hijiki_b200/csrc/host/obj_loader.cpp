// OBJ/MTL front-end: restates what `Scene::from_obj` (reference src/main.rs:413-530)
// gets out of `tobj::load_obj` (crate tobj 0.1.11, not vendored in the reference):
//   * `o`/`g` close the current model; a `usemtl` change with pending faces does too;
//   * faces are fan-triangulated (a,b,c),(a,c,d),...;
//   * a model's vertices are the unique (v,vt,vn) triples in first-use order;
//   * MTL materials keep file order; unknown keys (e.g. `Ke`) are kept as strings.
// Material class is chosen by NAME PREFIX (src/main.rs:432-458).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>

#include "host_scene.h"

namespace hjk {
namespace {

struct MtlMaterial {
  std::string name;
  float diffuse[3] = {0.f, 0.f, 0.f};
  std::map<std::string, std::string> unknown;
};

struct VertexIndices {
  int64_t v = -1, vt = -1, vn = -1;
  bool operator<(const VertexIndices& o) const {
    return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn);
  }
};

struct Mesh {
  std::vector<float> positions, normals, texcoords;
  std::vector<uint32_t> indices;
  int material_id = -1;
};

std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}

bool parse_floats(const std::vector<std::string>& words, size_t first, size_t n,
                  std::vector<float>& dst) {
  if (words.size() < first + n) return false;
  for (size_t i = 0; i < n; i++) {
    char* end = nullptr;
    float f = strtof(words[first + i].c_str(), &end);
    if (end == words[first + i].c_str()) return false;
    dst.push_back(f);
  }
  return true;
}

// one "v", "v/vt", "v//vn" or "v/vt/vn" token; negative = relative to the arrays so far
bool parse_face_vertex(const std::string& tok, size_t npos, size_t ntex, size_t nnorm,
                       VertexIndices& out) {
  int64_t vals[3] = {0, 0, 0};
  bool present[3] = {false, false, false};
  size_t field = 0, start = 0;
  for (size_t i = 0; i <= tok.size() && field < 3; i++) {
    if (i == tok.size() || tok[i] == '/') {
      if (i > start) {
        vals[field] = strtoll(tok.substr(start, i - start).c_str(), nullptr, 10);
        present[field] = true;
      }
      field++;
      start = i + 1;
    }
  }
  if (!present[0]) return false;
  const size_t counts[3] = {npos, ntex, nnorm};
  int64_t res[3] = {-1, -1, -1};
  for (int k = 0; k < 3; k++) {
    if (!present[k]) continue;
    res[k] = vals[k] < 0 ? (int64_t)counts[k] + vals[k] : vals[k] - 1;
    if (res[k] < 0) return false;
  }
  out.v = res[0];
  out.vt = res[1];
  out.vn = res[2];
  return true;
}

struct Face {
  std::vector<VertexIndices> verts;
};

Mesh export_faces(const std::vector<float>& pos, const std::vector<float>& tex,
                  const std::vector<float>& nrm, const std::vector<Face>& faces, int mat_id) {
  Mesh mesh;
  mesh.material_id = mat_id;
  std::map<VertexIndices, uint32_t> index_map;
  auto add_vertex = [&](const VertexIndices& vi) {
    auto it = index_map.find(vi);
    if (it != index_map.end()) {
      mesh.indices.push_back(it->second);
      return;
    }
    for (int k = 0; k < 3; k++) mesh.positions.push_back(pos[3 * vi.v + k]);
    if (!tex.empty() && vi.vt >= 0)
      for (int k = 0; k < 2; k++) mesh.texcoords.push_back(tex[2 * vi.vt + k]);
    if (!nrm.empty() && vi.vn >= 0)
      for (int k = 0; k < 3; k++) mesh.normals.push_back(nrm[3 * vi.vn + k]);
    uint32_t next = (uint32_t)index_map.size();
    mesh.indices.push_back(next);
    index_map.emplace(vi, next);
  };
  for (const Face& f : faces) {
    if (f.verts.size() < 3) continue;  // points / lines carry no surface
    for (size_t c = 2; c < f.verts.size(); c++) {
      add_vertex(f.verts[0]);
      add_vertex(f.verts[c - 1]);
      add_vertex(f.verts[c]);
    }
  }
  return mesh;
}

bool load_mtl(const std::string& path, std::vector<MtlMaterial>& mats,
              std::map<std::string, int>& mat_map) {
  std::ifstream in(path);
  if (!in) return false;
  std::string line;
  bool have = false;
  MtlMaterial cur;
  auto flush = [&]() {
    if (have) {
      mat_map[cur.name] = (int)mats.size();
      mats.push_back(cur);
    }
  };
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    auto words = split_ws(line);
    if (words.empty() || words[0][0] == '#') continue;
    const std::string& key = words[0];
    if (key == "newmtl") {
      flush();
      cur = MtlMaterial();
      cur.name = words.size() > 1 ? words[1] : "";
      have = true;
    } else if (key == "Kd") {
      std::vector<float> v;
      if (parse_floats(words, 1, 3, v)) {
        cur.diffuse[0] = v[0];
        cur.diffuse[1] = v[1];
        cur.diffuse[2] = v[2];
      }
    } else if (key == "Ka" || key == "Ks" || key == "Ns" || key == "Ni" || key == "d" ||
               key == "illum" || key == "map_Ka" || key == "map_Kd" || key == "map_Ks" ||
               key == "map_Ns" || key == "map_d" || key == "Tr") {
      // known to tobj, unused by the reference
    } else {
      // tobj: unknown_param[key] = rest of the line after the key, trimmed
      size_t p = line.find(key) + key.size();
      std::string rest = line.substr(p);
      size_t a = rest.find_first_not_of(" \t");
      size_t b = rest.find_last_not_of(" \t");
      cur.unknown[key] = a == std::string::npos ? "" : rest.substr(a, b - a + 1);
    }
  }
  flush();
  return true;
}

bool starts_with(const std::string& s, const char* prefix) {
  return s.compare(0, strlen(prefix), prefix) == 0;
}

}  // namespace

bool scene_from_obj(const std::string& path, Scene& scene, std::string& err) {
  std::ifstream in(path);
  if (!in) {
    err = "cannot open OBJ file: " + path;
    return false;
  }
  std::string dir;
  size_t slash = path.find_last_of('/');
  if (slash != std::string::npos) dir = path.substr(0, slash + 1);

  std::vector<Mesh> models;
  std::vector<MtlMaterial> mtl;
  std::map<std::string, int> mat_map;
  std::vector<float> pos, tex, nrm;
  std::vector<Face> faces;
  int mat_id = -1;

  std::string line;
  size_t line_no = 0;
  while (std::getline(in, line)) {
    line_no++;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    auto words = split_ws(line);
    if (words.empty() || words[0] == "#" || words[0][0] == '#') continue;
    const std::string& key = words[0];
    if (key == "v") {
      if (!parse_floats(words, 1, 3, pos)) {
        err = "bad position at line " + std::to_string(line_no);
        return false;
      }
    } else if (key == "vt") {
      if (!parse_floats(words, 1, 2, tex)) {
        err = "bad texcoord at line " + std::to_string(line_no);
        return false;
      }
    } else if (key == "vn") {
      if (!parse_floats(words, 1, 3, nrm)) {
        err = "bad normal at line " + std::to_string(line_no);
        return false;
      }
    } else if (key == "f" || key == "l") {
      Face f;
      for (size_t i = 1; i < words.size(); i++) {
        VertexIndices vi;
        if (!parse_face_vertex(words[i], pos.size() / 3, tex.size() / 2, nrm.size() / 3, vi) ||
            (size_t)vi.v >= pos.size() / 3 || (vi.vt >= 0 && (size_t)vi.vt >= tex.size() / 2) ||
            (vi.vn >= 0 && (size_t)vi.vn >= nrm.size() / 3)) {
          err = "bad face at line " + std::to_string(line_no);
          return false;
        }
        f.verts.push_back(vi);
      }
      faces.push_back(std::move(f));
    } else if (key == "o" || key == "g") {
      if (!faces.empty()) {
        models.push_back(export_faces(pos, tex, nrm, faces, mat_id));
        faces.clear();
      }
    } else if (key == "mtllib") {
      if (words.size() > 1 && !load_mtl(dir + words[1], mtl, mat_map)) {
        err = "cannot open MTL file: " + dir + words[1];
        return false;
      }
    } else if (key == "usemtl") {
      if (words.size() > 1) {
        auto it = mat_map.find(words[1]);
        int new_mat = it == mat_map.end() ? -1 : it->second;
        if (new_mat != mat_id && !faces.empty()) {
          models.push_back(export_faces(pos, tex, nrm, faces, mat_id));
          faces.clear();
        }
        mat_id = new_mat;
      }
    }
  }
  models.push_back(export_faces(pos, tex, nrm, faces, mat_id));

  // ---- Scene::from_obj proper (src/main.rs:417-529)
  scene = Scene();
  // `-1.45f32.to_radians()`: f32 multiply by (PI_f32 / 180), negated
  const float pi_f = 3.14159274101257324f;
  float angle = -(1.45f * (pi_f / 180.0f));
  scene.camera = HjkCamera{};
  scene.camera.position[0] = 0.f;
  scene.camera.position[1] = 0.91f;
  scene.camera.position[2] = 5.41f;
  scene.camera.position[3] = 0.f;
  scene.camera.rotation[0] = sinf(0.5f * angle);
  scene.camera.rotation[1] = 0.f;
  scene.camera.rotation[2] = 0.f;
  scene.camera.rotation[3] = cosf(0.5f * angle);
  scene.camera.fov = 27.7f;

  for (const MtlMaterial& m : mtl) {
    Material mat;
    if (starts_with(m.name, "light")) {
      auto it = m.unknown.find("Ke");
      if (it == m.unknown.end()) {
        err = "light material '" + m.name + "' has no Ke";
        return false;
      }
      std::vector<float> ke;
      if (!parse_floats(split_ws(it->second), 0, 3, ke)) {
        err = "light material '" + m.name + "' has a malformed Ke";
        return false;
      }
      mat.tag = HJK_MAT_EMISSIVE;
      mat.color = HjkColor16{{ke[0], ke[1], ke[2]}, 0.f};
    } else if (starts_with(m.name, "glass")) {
      mat.tag = HJK_MAT_DIELECTRIC;
      mat.dielectric = HjkDielectric{{0.f, 0.f, 0.f, 1.5f}};  // DielectricMaterial::clear(1.5)
    } else if (starts_with(m.name, "mirror")) {
      mat.tag = HJK_MAT_MIRROR;
    } else {
      mat.tag = HJK_MAT_DIFFUSE;
      mat.color = HjkColor16{{m.diffuse[0], m.diffuse[1], m.diffuse[2]}, 0.f};
    }
    scene.materials.push_back(mat);
  }

  for (const Mesh& mesh : models) {
    uint32_t vertex_offset = (uint32_t)scene.vertices.size();
    size_t nv = mesh.positions.size() / 3;
    for (size_t i = 0; i < nv; i++) {
      HjkVertex v{};
      v.pos[0] = mesh.positions[3 * i];
      v.pos[1] = mesh.positions[3 * i + 1];
      v.pos[2] = mesh.positions[3 * i + 2];
      if (mesh.texcoords.size() >= 2 * (i + 1)) {  // `.get(2*i..2*(i+1)).unwrap_or(&[0.,0.])`
        v.u = mesh.texcoords[2 * i];
        v.v = mesh.texcoords[2 * i + 1];
      }
      if (mesh.normals.size() < 3 * (i + 1)) {  // `.unwrap()` on a missing normal
        err = "model has a vertex without a normal (reference panics here, src/main.rs:467)";
        return false;
      }
      v.normal[0] = mesh.normals[3 * i];
      v.normal[1] = mesh.normals[3 * i + 1];
      v.normal[2] = mesh.normals[3 * i + 2];
      scene.vertices.push_back(v);
    }
    if (mesh.material_id < 0) continue;  // `_ => continue` (src/main.rs:476-479)
    for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
      Shape s;
      s.kind = ShapeKind::Triangle;
      s.tri = {mesh.indices[t] + vertex_offset, mesh.indices[t + 1] + vertex_offset,
               mesh.indices[t + 2] + vertex_offset};
      scene.objects.emplace_back(s, (size_t)mesh.material_id);
      // the quad-recovery block after the `continue` (src/main.rs:489-525) is dead code
    }
  }
  return true;
}

void put_cbox_spheres(Scene& scene) {
  Material mirror;
  mirror.tag = HJK_MAT_MIRROR;
  scene.materials.push_back(mirror);
  Material cb;
  cb.tag = HJK_MAT_DIFFUSECBOARD;
  cb.cboard = HjkDiffuseCB{{1.0f, 0.4f, 0.7f}, 0.1f, {0.4f, 0.7f, 1.0f}, 0.2f};
  scene.materials.push_back(cb);
  Shape a;
  a.kind = ShapeKind::Sphere;
  a.sphere = HjkSphere{{-0.421400f, 0.332100f, -0.280000f}, 0.3263f};
  scene.objects.emplace_back(a, scene.materials.size() - 2);
  Shape b;
  b.kind = ShapeKind::Sphere;
  b.sphere = HjkSphere{{0.445800f, 0.332100f, 0.376700f}, 0.3263f};
  scene.objects.emplace_back(b, scene.materials.size() - 1);
}

}  // namespace hjk
