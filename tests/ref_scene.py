"""An INDEPENDENT restatement, in Python, of the reference's scene front-end — test infrastructure.

    Scene::from_obj       reference src/main.rs:413-530  (over tobj::load_obj, crate tobj 0.1.x, not vendored)
    --put-cbox-spheres    reference src/main.rs:1463-1483
    Scene::compile        reference src/main.rs:172-358  (everything but the `bvh` binding)

The product's loader is C++ (hijiki_b200/csrc/host/obj_loader.cpp, scene_compile.cpp); the oracle takes its
scene arrays from whoever calls it.  So that at least one parity chain does not hand BOTH sides arrays made by
the product's own loader, this module builds the twelve CompiledScene arrays from the OBJ/MTL text with no code
shared with the product (pure Python + numpy, written from the Rust source), and tests/test_ref_scene.py holds the
product's loader to it byte for byte — a renumbered shape, a reordered material or a different emitter table in
scene_compile.cpp fails there.

tobj behaviour restated (tobj 0.1.x `load_obj`): `o`/`g` start a new model; faces are fan-triangulated; a model's
vertices are the distinct (v, vt, vn) index triples in first-use order; MTL materials keep file order, `Kd` is the
diffuse colour, keys tobj does not know (`Ke`) are kept as strings.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from hijiki_b200 import _abi

F = np.float32
TAG = {"diffuse": 0, "diffusecb": 1, "mirror": 2, "dielectric": 3, "emissive": 4}  # src/main.rs:34-43


def _parse_mtl(path):
    mats = []
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w or w[0].startswith("#"):
                continue
            if w[0] == "newmtl":
                mats.append({"name": " ".join(w[1:]), "Kd": [F(0), F(0), F(0)], "unknown": {}})
            elif mats:
                if w[0] == "Kd":
                    mats[-1]["Kd"] = [F(x) for x in w[1:4]]
                elif w[0] not in ("Ka", "Ks", "Ns", "Ni", "d", "illum", "map_Ka", "map_Kd", "map_Ks", "map_Ns",
                                  "map_d", "Tr"):
                    mats[-1]["unknown"][w[0]] = " ".join(w[1:])
    return mats


def _parse_obj(path):
    """-> (models, mtl materials); a model = dict(positions, normals, texcoords, indices, material_id)."""
    pos, nrm, tex = [], [], []
    models, mtl = [], []
    faces, cur_mat = [], None
    mat_index = {}

    def flush():
        nonlocal faces
        if not faces:
            return
        m = {"positions": [], "normals": [], "texcoords": [], "indices": [], "material_id": cur_mat}
        seen = {}
        for face in faces:
            for c in range(2, len(face)):
                for key in (face[0], face[c - 1], face[c]):
                    k = seen.get(key)
                    if k is None:
                        k = len(seen)
                        seen[key] = k
                        v, vt, vn = key
                        m["positions"] += pos[v]
                        if vt is not None and tex:
                            m["texcoords"] += tex[vt]
                        if vn is not None and nrm:
                            m["normals"] += nrm[vn]
                    m["indices"].append(k)
        models.append(m)
        faces = []

    def index(tok, n):
        i = int(tok)
        return i - 1 if i > 0 else n + i

    with open(path) as f:
        for line in f:
            w = line.split()
            if not w or w[0].startswith("#"):
                continue
            if w[0] == "v":
                pos.append([F(x) for x in w[1:4]])
            elif w[0] == "vn":
                nrm.append([F(x) for x in w[1:4]])
            elif w[0] == "vt":
                tex.append([F(x) for x in w[1:3]])
            elif w[0] == "f":
                face = []
                for tok in w[1:]:
                    parts = tok.split("/")
                    v = index(parts[0], len(pos))
                    vt = index(parts[1], len(tex)) if len(parts) > 1 and parts[1] else None
                    vn = index(parts[2], len(nrm)) if len(parts) > 2 and parts[2] else None
                    face.append((v, vt, vn))
                if len(face) >= 3:
                    faces.append(face)
            elif w[0] in ("o", "g"):
                flush()
            elif w[0] == "usemtl":
                name = " ".join(w[1:])
                new = mat_index.get(name)
                if new != cur_mat and faces:
                    flush()
                cur_mat = new
            elif w[0] == "mtllib":
                for m in _parse_mtl(os.path.join(os.path.dirname(path), " ".join(w[1:]))):
                    mat_index.setdefault(m["name"], len(mtl))
                    mtl.append(m)
    flush()
    return models, mtl


class RefScene:
    """The CompiledScene arrays (binding order of src/main.rs:314-327; no `bvh`) built by this module."""

    def __init__(self, obj_path: str, put_cbox_spheres: bool = False):
        models, mtl = _parse_obj(obj_path)
        # ---- Scene::from_obj, src/main.rs:413-530
        angle = F(-1.45) * F(np.pi / 180.0)  # f32::to_radians = x * (PI / 180)
        half = F(0.5) * angle
        camera = ((F(0.0), F(0.91), F(5.41), F(0.0)), (F(np.sin(half)), F(0.0), F(0.0), F(np.cos(half))), F(27.7))
        materials = []  # (class, payload)
        for m in mtl:  # src/main.rs:432-458: the class is chosen by name prefix
            if m["name"].startswith("light"):
                materials.append(("emissive", [F(x) for x in m["unknown"]["Ke"].split(" ")][:3]))
            elif m["name"].startswith("glass"):
                materials.append(("dielectric", [F(0), F(0), F(0), F(1.0) / F(1.5)]))
            elif m["name"].startswith("mirror"):
                materials.append(("mirror", None))
            else:
                materials.append(("diffuse", m["Kd"]))
        vertices, objects = [], []  # objects: (kind, data, material)
        for model in models:
            offset = len(vertices)
            p, n, t = model["positions"], model["normals"], model["texcoords"]
            for i in range(len(p) // 3):
                uv = t[2 * i:2 * i + 2] if len(t) >= 2 * i + 2 else [F(0), F(0)]
                nn = n[3 * i:3 * i + 3]
                assert len(nn) == 3, "the reference unwraps the normal (src/main.rs:466)"
                vertices.append([p[3 * i], p[3 * i + 1], p[3 * i + 2], uv[0], nn[0], nn[1], nn[2], uv[1]])
            if model["material_id"] is None:
                continue
            idx = model["indices"]
            for k in range(0, len(idx), 3):
                objects.append(("triangle", [idx[k] + offset, idx[k + 1] + offset, idx[k + 2] + offset],
                                model["material_id"]))
        if put_cbox_spheres:  # src/main.rs:1463-1483
            materials.append(("mirror", None))
            materials.append(("diffusecb", [F(1.0), F(0.4), F(0.7), F(0.1), F(0.4), F(0.7), F(1.0), F(0.2)]))
            objects.append(("sphere", [F(-0.421400), F(0.332100), F(-0.280000), F(0.3263)], len(materials) - 2))
            objects.append(("sphere", [F(0.445800), F(0.332100), F(0.376700), F(0.3263)], len(materials) - 1))
        # ---- Scene::compile, src/main.rs:172-358
        spheres = [(d, m) for k, d, m in objects if k == "sphere"]
        quads = [(d, m) for k, d, m in objects if k == "quad"]
        triangles = [(d, m) for k, d, m in objects if k == "triangle"]
        pools = {"diffuse": [], "diffusecb": [], "dielectric": [], "emissive": []}
        reprs = []
        for cls, payload in materials:  # src/main.rs:258-283
            if cls == "mirror":
                ix = 0
            else:
                pools[cls].append(payload)
                ix = len(pools[cls]) - 1
            reprs.append((TAG[cls] << 24) + ix)
        mat_words = [reprs[m] for _, m in spheres] + [reprs[m] for _, m in quads] + [reprs[m] for _, m in triangles]
        em_shapes = [i for i, w in enumerate(mat_words) if w >> 24 == TAG["emissive"]]
        emitters = np.zeros((len(em_shapes), 4), F)
        if em_shapes:  # src/main.rs:296-312: pdf = 1 / n, cdf accumulated in f32
            pdf = F(1.0) / F(len(em_shapes))
            cdf = F(0.0)
            for k, sh in enumerate(em_shapes):
                cdf = F(cdf + pdf)
                emitters[k, 1], emitters[k, 2] = pdf, cdf
            emitters.view(np.uint32)[:, 0] = em_shapes
        pad = lambda rows, n: np.array([list(r) + [F(0)] * (n - len(r)) for r in rows], F).reshape(-1, n)
        self.arrays = {
            "spheres": np.array([d for d, _ in spheres], F).reshape(-1, 4),
            "quads": np.zeros((0, 12), F),
            "triangles": np.array([d for d, _ in triangles], np.uint32).reshape(-1, 3),
            "vertices": np.array(vertices, F).reshape(-1, 8),
            "materials": np.array(mat_words, np.uint32),
            "emitters": emitters,
            "diffuse": pad(pools["diffuse"], 4),
            "diffusecb": np.array(pools["diffusecb"], F).reshape(-1, 8),
            "dielectric": np.array(pools["dielectric"], F).reshape(-1, 4),
            "emissive": pad(pools["emissive"], 4),
        }
        assert not quads  # from_obj's quad recovery is dead code (`continue`, src/main.rs:487)
        self.info_struct = _abi.HjkSceneInfo()
        for k in range(4):
            self.info_struct.camera.position[k] = camera[0][k]
            self.info_struct.camera.rotation[k] = camera[1][k]
        self.info_struct.camera.fov = camera[2]
        self.info_struct.num_spheres = len(spheres)
        self.info_struct.num_quads = 0
        self.info_struct.num_triangles = len(triangles)
        self.info_struct.num_emitters = len(em_shapes)
        self.view = _abi.HjkScene()
        self.view.scene = _abi.HjkArray(C.addressof(self.info_struct), 1)
        self.view.bvh = _abi.HjkArray(None, 0)
        for name, a in self.arrays.items():
            a = np.ascontiguousarray(a)
            self.arrays[name] = a
            setattr(self.view, name, _abi.HjkArray(a.ctypes.data if a.size else None, len(a)))

    @property
    def info(self):
        return self.info_struct

    def array(self, name):
        return self.arrays[name]
