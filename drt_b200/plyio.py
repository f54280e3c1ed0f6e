"""Binary/ASCII PLY triangle-mesh I/O (the subset trimesh.load / mesh.export covers for
DiffRender.py:303-309 and optim.py:50,226).  trimesh is not a dependency of this package.

All 16 meshes of the reference are binary little-endian, float x/y/z (+ optional extra float
properties such as `quality`), faces as `list uchar int vertex_indices`.
"""
import numpy as np

_PLY_DT = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def read_ply(path):
    """-> (vertices float64 [V,3], faces int64 [F,3]).  Extra vertex properties are skipped."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []  # (name, count, [(prop_name, dtype | ('list', cnt_dt, item_dt))])
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], ("list", _PLY_DT[tok[2]], _PLY_DT[tok[3]])))
                else:
                    elements[-1][2].append((tok[2], _PLY_DT[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
        end = ">" if fmt == "binary_big_endian" else "<"
        verts = faces = None
        for name, count, props in elements:
            has_list = any(isinstance(dt, tuple) for _, dt in props)
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(count)]
                if name == "vertex":
                    names = [p for p, _ in props]
                    ix = [names.index(c) for c in ("x", "y", "z")]
                    verts = np.array([[float(r[i]) for i in ix] for r in rows], dtype=np.float64).reshape(-1, 3)
                elif name == "face":
                    for r in rows:
                        if int(r[0]) != 3:
                            raise ValueError(f"{path}: non-triangular face")
                    faces = np.array([[int(v) for v in r[1:4]] for r in rows], dtype=np.int64).reshape(-1, 3)
                continue
            if not has_list:
                dt = np.dtype([(p, end + t) for p, t in props])
                data = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                if name == "vertex":
                    verts = np.stack([data["x"], data["y"], data["z"]], axis=1).astype(np.float64)
            else:
                if name != "face" or len(props) != 1:
                    raise ValueError(f"{path}: unsupported list element {name!r}")
                _, (_, cdt, idt) = props[0]
                dt = np.dtype([("n", end + cdt), ("v", end + idt, (3,))])
                data = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                if count and not (data["n"] == 3).all():
                    raise ValueError(f"{path}: non-triangular face")
                faces = data["v"].astype(np.int64)
        if verts is None or faces is None:
            raise ValueError(f"{path}: needs vertex and face elements")
        return verts, faces


def write_ply(path, vertices, faces):
    """Binary little-endian PLY, float32 vertices + uchar/int32 faces -- the layout of the
    reference's own meshes, readable by MeshLab (optim.py:46-52)."""
    v = np.ascontiguousarray(vertices, dtype="<f4").reshape(-1, 3)
    fa = np.ascontiguousarray(faces, dtype="<i4").reshape(-1, 3)
    rec = np.empty(len(fa), dtype=[("n", "u1"), ("v", "<i4", (3,))])
    rec["n"] = 3
    rec["v"] = fa
    header = (
        "ply\nformat binary_little_endian 1.0\ncomment drt_b200\n"
        f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
        f"element face {len(fa)}\nproperty list uchar int vertex_indices\nend_header\n"
    )
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(v.tobytes())
        f.write(rec.tobytes())
