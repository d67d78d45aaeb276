"""Init-time inputs of the path, generated host-side with numpy exactly as the reference's topology
components produce them (ordering matters for parity: the gather sums follow element order).

  RegularGridTopology   Sofa/Component/Topology/Container/Grid/src/sofa/component/topology/container/grid/
                        RegularGridTopology.cpp:124-168, GridTopology.cpp:352-398
  Hexa2TetraTopologicalMapping   Sofa/Component/Topology/Mapping/src/sofa/component/topology/mapping/
                        Hexa2TetraTopologicalMapping.cpp:104-196
  TetrahedronFEMForceField::init's own tessellation   .../fem/elastic/TetrahedronFEMForceField.inl:1317-1371
  BoxROI                Sofa/Component/Engine/Select/src/sofa/component/engine/select/BoxROI.inl:189-201
  DiagonalMass lumping  Sofa/Component/Mass/src/sofa/component/mass/DiagonalMass.inl:1061-1110,1218-1258
  Gmsh v1 reader        Sofa/Component/IO/Mesh/src/sofa/component/io/mesh/MeshGmshLoader.cpp
"""
import numpy as np

# Sofa/framework/Geometry/src/sofa/geometry/Hexahedron.h:67-88
_X_EDGES = ((0, 1), (4, 5), (3, 2), (7, 6))
_Y_EDGES = ((4, 7), (5, 6), (1, 2), (0, 3))
_Z_EDGES = ((4, 0), (5, 1), (6, 2), (7, 3))
_NON_SWAPPED = np.array([[0, 5, 1, 6], [0, 1, 3, 6], [1, 3, 6, 2], [6, 3, 0, 7], [6, 7, 0, 5], [7, 5, 4, 0]])
_SWAPPED = np.array([[0, 5, 6, 1], [0, 1, 6, 3], [1, 3, 2, 6], [6, 3, 7, 0], [6, 7, 5, 0], [7, 5, 0, 4]])


def regular_grid(n, mn, mx):
    """RegularGridTopology(n, min, max) -> (positions float64 [N,3], hexahedra uint32 [H,8])."""
    nx, ny, nz = (int(v) for v in n)
    mn = np.asarray(mn, np.float64); mx = np.asarray(mx, np.float64)
    p0 = mn.copy(); d = np.empty(3)
    for c, m in enumerate((nx - 1, ny - 1, nz - 1)):
        if m > 0:
            d[c] = (mx[c] - mn[c]) / m
        else:
            d[c] = mx[c] - mn[c]
            if c < 2:
                p0[c] = (mx[c] + mn[c]) / 2
    i = np.arange(nx, dtype=np.float64); j = np.arange(ny, dtype=np.float64); k = np.arange(nz, dtype=np.float64)
    pos = np.empty((nz, ny, nx, 3), np.float64)
    pos[..., 0] = (p0[0] + d[0] * i)[None, None, :]
    pos[..., 1] = (p0[1] + d[1] * j)[None, :, None]
    pos[..., 2] = (p0[2] + d[2] * k)[:, None, None]
    pos = pos.reshape(-1, 3)
    if min(nx, ny, nz) < 2:
        return pos, np.zeros((0, 8), np.uint32)
    z, y, x = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    x = x.ravel(); y = y.ravel(); z = z.ravel()

    def P(a, b, c):
        return nx * (ny * c + b) + a
    hexas = np.stack([P(x, y, z), P(x + 1, y, z), P(x + 1, y + 1, z), P(x, y + 1, z),
                      P(x, y, z + 1), P(x + 1, y, z + 1), P(x + 1, y + 1, z + 1), P(x, y + 1, z + 1)], axis=1).astype(np.uint32)
    return pos, hexas


def hexas_to_tetras(hexas, n, mode="mapping"):
    """6 tetrahedra per hexahedron of a grid with n=(nx,ny,nz) points.
    mode: "mapping" (Hexa2TetraTopologicalMapping swapping=false), "mapping_swapping" (swapping=true),
          "forcefield" (the tessellation TetrahedronFEMForceField::init builds for a hexahedral topology)."""
    H = hexas.shape[0]
    nx, ny = int(n[0]) - 1, int(n[1]) - 1
    c = hexas.astype(np.uint32).copy()
    swapped = np.zeros(H, bool)
    if mode != "mapping":
        i = np.arange(H)
        for cond, edges in ((((i % nx) & 1) == 0, _X_EDGES), ((((i // nx) % ny) & 1) == 1, _Y_EDGES), (((i // (nx * ny)) & 1) == 1, _Z_EDGES)):
            for a, b in edges:
                tmp = c[cond, a].copy(); c[cond, a] = c[cond, b]; c[cond, b] = tmp
            swapped ^= cond
    tets = np.empty((H, 6, 4), np.uint32)
    use_sw = swapped if mode == "mapping_swapping" else np.zeros(H, bool)
    for t in range(6):
        for k in range(4):
            tets[:, t, k] = np.where(use_sw, c[np.arange(H), _SWAPPED[t, k]], c[np.arange(H), _NON_SWAPPED[t, k]])
    return tets.reshape(-1, 4)


def box_roi(positions, box):
    b = np.asarray(box, np.float64)
    return np.nonzero(np.all((positions >= b[:3]) & (positions <= b[3:]), axis=1))[0].astype(np.uint32)


def _tet_volume(p, t):
    a = p[t[:, 1]] - p[t[:, 0]]; b = p[t[:, 2]] - p[t[:, 0]]; c = p[t[:, 3]] - p[t[:, 0]]
    cx = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    cy = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    cz = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    r = cx * c[:, 0]; r = r + cy * c[:, 1]; r = r + cz * c[:, 2]
    return np.abs(r / p.dtype.type(6))


def diagonal_mass(positions, elems, dtype, mass_density=None, total_mass=None):
    """DiagonalMass vertexMass lumped from tetrahedra ([T,4]) or hexahedra ([H,8]) in the reference's Real arithmetic."""
    dt = np.dtype(dtype).type
    p = np.ascontiguousarray(positions, dtype)
    e = np.asarray(elems, np.int64)
    density = dt(1.0) if mass_density is None else dt(mass_density)
    if e.shape[1] == 4:
        vol = _tet_volume(p, e); share = dt(4.0)
    else:
        idx = ((0, 5, 1, 6), (0, 1, 3, 6), (1, 3, 6, 2), (6, 3, 0, 7), (6, 7, 0, 5), (7, 5, 4, 0))
        vol = None
        for q in idx:
            v = _tet_volume(p, e[:, q])
            vol = v if vol is None else vol + v
        share = dt(8.0)
    m_el = (density * vol) / share
    masses = np.zeros(p.shape[0], dtype)
    np.add.at(masses, e.ravel(), np.repeat(m_el, e.shape[1]))  # sequential, element order then corner order
    if total_mass is not None:
        # initFromTotalMass: density = totalMass / sum (sum accumulated in the same order, in Real)
        flat = np.repeat(m_el, e.shape[1]).astype(dtype)
        s = np.cumsum(flat, dtype=dtype)[-1] if flat.size else dt(0)  # cumsum == strictly sequential `total_mass += mass` in Real
        dens = dt(1.0) if s < np.finfo(dtype).eps else dt(dt(total_mass) / s)
        masses = masses * dens
    return masses


def tetra_edges(tets):
    """Edge array of a tetrahedral topology in the order TetrahedronSetTopologyContainer::createEdgeSetArray builds it (first appearance
    over the tetrahedra, local edges {0,1},{0,2},{0,3},{1,2},{1,3},{2,3}, vertices sorted) and the 6 edge ids of every tetrahedron."""
    t = np.asarray(tets, np.int64).reshape(-1, 4)
    loc = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])
    pairs = np.sort(t[:, loc], axis=2).reshape(-1, 2)            # tetra-major, local-edge order
    nmax = int(t.max()) + 1 if t.size else 1
    key = pairs[:, 0] * nmax + pairs[:, 1]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                       # unique keys ranked by first appearance
    rank = np.empty_like(order); rank[order] = np.arange(order.size)
    edges = pairs[first[order]].astype(np.uint32)
    return edges, rank[inv].reshape(-1, 6).astype(np.uint32)


def mesh_matrix_mass(positions, tets, dtype, mass_density=1.0, lumping=False):
    """MeshMatrixMass on tetrahedra (MeshMatrixMass.inl:547-665,1476-1490) in the reference's Real arithmetic:
    returns (vertexMass, edges, edgeMass, massLumpingCoeff)."""
    dt = np.dtype(dtype).type
    p = np.ascontiguousarray(positions, dtype)
    t = np.asarray(tets, np.int64).reshape(-1, 4)
    vol = _tet_volume(p, t)
    edges, eit = tetra_edges(t)
    vm = np.zeros(p.shape[0], dtype)
    np.add.at(vm, t.ravel(), np.repeat((dt(mass_density) * vol) / dt(10.0), 4))        # sequential: tetra order, then corner order
    em = np.zeros(edges.shape[0], dtype)
    if not lumping:
        np.add.at(em, eit.astype(np.int64).ravel(), np.repeat((dt(mass_density) * vol) / dt(20.0), 6))
    return vm, edges, em, 2.5


def read_gmsh_v1(path):
    """Gmsh file format 1.0 ($NOD / $ELM) -> (positions float64 [N,3], tetrahedra uint32 [T,4], hexahedra uint32 [H,8])."""
    with open(path) as fh:
        tok = fh.read().split()
    i = tok.index("$NOD") + 1
    n = int(tok[i]); i += 1
    ids = {}
    pos = np.empty((n, 3), np.float64)
    for k in range(n):
        ids[int(tok[i])] = k
        pos[k] = [float(tok[i + 1]), float(tok[i + 2]), float(tok[i + 3])]
        i += 4
    i = tok.index("$ELM") + 1
    ne = int(tok[i]); i += 1
    tets, hexas = [], []
    for _ in range(ne):
        etype, nnodes = int(tok[i + 1]), int(tok[i + 4])
        nodes = [ids[int(v)] for v in tok[i + 5:i + 5 + nnodes]]
        if etype == 4:
            tets.append(nodes)
        elif etype == 5:
            hexas.append(nodes)
        i += 5 + nnodes
    return pos, np.array(tets, np.uint32).reshape(-1, 4), np.array(hexas, np.uint32).reshape(-1, 8)


def read_gmsh_v2(path):
    """Gmsh file format 2.x, ASCII ($MeshFormat 2.* / $Nodes / $Elements: "id type ntags tags... nodes...") -> (positions,
    tetrahedra, hexahedra), the subset MeshGmshLoader feeds this path (Sofa/Component/IO/Mesh/src/sofa/component/io/mesh/MeshGmshLoader.cpp:60-100,
    element types 4 = tetrahedron, 5 = hexahedron)."""
    with open(path) as fh:
        tok = fh.read().split()
    i = tok.index("$MeshFormat") + 1
    if int(float(tok[i])) != 2 or int(tok[i + 1]) != 0:
        raise ValueError(f"{path}: only ASCII MSH 2.x is read here (header: {tok[i]} {tok[i + 1]})")
    i = tok.index("$Nodes") + 1
    n = int(tok[i]); i += 1
    ids = {}
    pos = np.empty((n, 3), np.float64)
    for k in range(n):
        ids[int(tok[i])] = k
        pos[k] = [float(tok[i + 1]), float(tok[i + 2]), float(tok[i + 3])]
        i += 4
    i = tok.index("$Elements") + 1
    ne = int(tok[i]); i += 1
    nnodes_of = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 6: 6, 7: 5, 15: 1}
    tets, hexas = [], []
    for _ in range(ne):
        etype, ntags = int(tok[i + 1]), int(tok[i + 2])
        if etype not in nnodes_of:
            raise ValueError(f"{path}: element type {etype} is not supported")
        nn = nnodes_of[etype]
        nodes = [ids[int(v)] for v in tok[i + 3 + ntags:i + 3 + ntags + nn]]
        if etype == 4:
            tets.append(nodes)
        elif etype == 5:
            hexas.append(nodes)
        i += 3 + ntags + nn
    return pos, np.array(tets, np.uint32).reshape(-1, 4), np.array(hexas, np.uint32).reshape(-1, 8)


def read_gmsh(path):
    """Gmsh .msh, format 1.0 or 2.x (ASCII), chosen by the first keyword like MeshGmshLoader does."""
    with open(path) as fh:
        first = fh.readline().strip()
    if first.startswith("$MeshFormat"):
        return read_gmsh_v2(path)
    if first.startswith("$NOD"):
        return read_gmsh_v1(path)
    raise ValueError(f"{path}: neither $MeshFormat nor $NOD at the top: not a registered MSH format")


# ---------------------------------------------------------------------------------------------------
# WriteState / ReadState text dumps (Sofa/Component/Playback/src/sofa/component/playback/WriteState.inl:350-381,
# ReadState.inl:219-262): one block per exported time, "T= <time>" then "  X= x0 y0 z0 x1 ...", "  V= ...", optionally "  F= ...",
# "  X0= ...".  The reference writes with the stream's default precision (6 significant digits); `precision=17` keeps doubles exact.
# ---------------------------------------------------------------------------------------------------
def read_vtk_legacy(path):
    """Legacy VTK file ("# vtk DataFile Version x.y"), DATASET UNSTRUCTURED_GRID, ASCII or BINARY (big-endian), the volumetric subset
    MeshVTKLoader feeds this path (Sofa/Component/IO/Mesh/src/sofa/component/io/mesh/MeshVTKLoader.cpp: LegacyVTKReader; cell types
    10 = VTK_TETRA, 12 = VTK_HEXAHEDRON, same corner order as SOFA's Tetrahedron / Hexahedron) -> (positions, tetrahedra, hexahedra).
    Other cell types are skipped; POINT_DATA / CELL_DATA are ignored."""
    with open(path, "rb") as fh:
        raw = fh.read()
    pos_in_file = 0

    def line():
        nonlocal pos_in_file
        while True:
            j = raw.find(b"\n", pos_in_file)
            if j < 0:
                j = len(raw)
            ln = raw[pos_in_file:j].decode("ascii", "replace").strip()
            pos_in_file = min(j + 1, len(raw))
            if ln or pos_in_file >= len(raw):
                return ln

    if not line().lower().startswith("# vtk datafile"):
        raise ValueError(f"{path}: not a legacy VTK file")
    line()                                   # title
    binary = line().upper() == "BINARY"
    ds = line().upper().split()
    if ds[:2] != ["DATASET", "UNSTRUCTURED_GRID"]:
        raise ValueError(f"{path}: only DATASET UNSTRUCTURED_GRID is read here (found {' '.join(ds)})")
    vtk_types = {"float": ">f4", "double": ">f8", "int": ">i4", "unsigned_int": ">u4", "long": ">i8", "vtktypeint64": ">i8", "vtktypeint32": ">i4",
                 "unsigned_char": ">u1", "char": ">i1", "short": ">i2", "unsigned_short": ">u2", "unsigned_long": ">u8"}

    def numbers(count, vtype):
        nonlocal pos_in_file
        if binary:
            dt = np.dtype(vtk_types[vtype])
            a = np.frombuffer(raw, dt, count, pos_in_file)
            pos_in_file += count * dt.itemsize
            return a
        out = []
        while len(out) < count:
            out.extend(line().split())
        return np.array(out[:count], np.float64 if vtype in ("float", "double") else np.int64)

    pos = np.zeros((0, 3)); cells = None; ctypes_ = None; n_cells = 0
    while pos_in_file < len(raw):
        w = line().split()
        if not w:
            continue
        key = w[0].upper()
        if key == "POINTS":
            pos = numbers(3 * int(w[1]), w[2].lower()).astype(np.float64).reshape(-1, 3)
        elif key == "CELLS":
            n_cells = int(w[1]); cells = numbers(int(w[2]), "int").astype(np.int64)
        elif key == "CELL_TYPES":
            ctypes_ = numbers(int(w[1]), "int").astype(np.int64)
        elif key in ("POINT_DATA", "CELL_DATA"):
            break
    if cells is None or ctypes_ is None:
        raise ValueError(f"{path}: CELLS / CELL_TYPES missing")
    tets, hexas = [], []
    i = 0
    for c in range(n_cells):
        k = int(cells[i]); nodes = cells[i + 1:i + 1 + k]; i += 1 + k
        if ctypes_[c] == 10 and k == 4:
            tets.append(nodes)
        elif ctypes_[c] == 12 and k == 8:
            hexas.append(nodes)
    return pos, np.array(tets, np.uint32).reshape(-1, 4), np.array(hexas, np.uint32).reshape(-1, 8)


def write_state(path, frames, precision=6):
    """frames: iterable of dicts {"T": time, "X": [n,3], "V": [n,3], ...} (keys other than T are optional)."""
    fmt = f"%.{int(precision)}g"
    with open(path, "w") as fh:
        for fr in frames:
            fh.write("T= " + (fmt % float(fr["T"])) + "\n")
            for key in ("X", "X0", "V", "F"):
                if key in fr and fr[key] is not None:
                    a = np.asarray(fr[key], np.float64).reshape(-1)
                    fh.write(f"  {key}= " + " ".join(fmt % v for v in a) + "\n")


def read_state(path):
    """Inverse of write_state / of SOFA's WriteState: list of {"T": float, "X": [n,3], "V": [n,3], ...}."""
    frames = []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            cmd = tok[0]
            if cmd == "T=":
                frames.append({"T": float(tok[1])})
            elif cmd in ("X=", "X0=", "V=", "F=") and frames:
                a = np.array([float(v) for v in tok[1:]], np.float64)
                if a.size % 3:
                    raise ValueError(f"{path}: {cmd} holds {a.size} values, not a multiple of 3")
                frames[-1][cmd[:-1]] = a.reshape(-1, 3)
    return frames
