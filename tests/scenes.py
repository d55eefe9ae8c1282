"""Small deterministic scenes shared by the CPU and GPU tests (edge cases the reference's shaders have to get right)."""
import numpy as np

from vk_gltf_viewer_b200.scene import Camera, Scene


def grid_mesh(nu, nv, fn, flip=False):
    """lofted grid like host/procedural.cpp: (nu+1)*(nv+1) vertices, 2*nu*nv triangles"""
    us, vs = np.meshgrid(np.linspace(0, 1, nu + 1, dtype=np.float32), np.linspace(0, 1, nv + 1, dtype=np.float32))
    pos = np.stack(fn(us, vs), axis=-1).reshape(-1, 3).astype(np.float32)
    idx = []
    for j in range(nv):
        for i in range(nu):
            a = j * (nu + 1) + i
            b, c, d = a + 1, a + nu + 1, a + nu + 2
            idx += ([a, c, b, b, c, d] if flip else [a, b, c, b, d, c])
    return pos, np.asarray(idx, np.uint32)


def single_triangle(double_sided=True):
    s = Scene.new()
    m = s.add_material(double_sided=double_sided)
    p = s.add_primitive([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], [0, 1, 2], m)
    s.add_node(p)
    return s.finalize()


def fullscreen_quad(z=-5.0, size=50.0):
    """two huge triangles far beyond the viewport: exercises the guard-band clipper and the cooperative path"""
    s = Scene.new()
    m = s.add_material(double_sided=True)
    p = s.add_primitive([[-size, -size, z], [size, -size, z], [size, size, z], [-size, size, z]], [0, 1, 2, 0, 2, 3], m)
    s.add_node(p)
    return s.finalize()


def ground_plane(n=24, extent=40.0, y=-1.0):
    """tessellated floor passing under (and behind) the camera: near-plane clipping + w<=0 vertices"""
    pos, idx = grid_mesh(n, n, lambda u, v: ((u * 2 - 1) * extent, np.full_like(u, y), (v * 2 - 1) * extent), flip=True)
    s = Scene.new()
    m = s.add_material(double_sided=True)
    p = s.add_primitive(pos, idx, m)
    s.add_node(p)
    return s.finalize()


def coplanar_overlap():
    """two identical quads in the same plane from two nodes: every covered pixel is an exact depth tie"""
    s = Scene.new()
    m = s.add_material(double_sided=True)
    pos, idx = grid_mesh(6, 6, lambda u, v: (u * 2 - 1, v * 2 - 1, np.zeros_like(u)))
    p = s.add_primitive(pos, idx, m)
    s.add_node(p)
    s.add_node(p)
    return s.finalize()


def occluder_and_hidden(n_hidden=6):
    """a big wall in front of a few small tessellated cards: second-frame HiZ must reject the cards"""
    s = Scene.new()
    m = s.add_material(double_sided=True)
    wall, widx = grid_mesh(16, 16, lambda u, v: ((u * 2 - 1) * 6, (v * 2 - 1) * 6, np.zeros_like(u)))
    pw = s.add_primitive(wall, widx, m)
    s.add_node(pw, translation=(0, 0, 0))
    card, cidx = grid_mesh(4, 4, lambda u, v: ((u * 2 - 1) * 0.3, (v * 2 - 1) * 0.3, np.zeros_like(u)))
    pc = s.add_primitive(card, cidx, m)
    rng = np.random.default_rng(7)
    for _ in range(n_hidden):
        x, y = rng.uniform(-2, 2, 2)
        s.add_node(pc, translation=(float(x), float(y), -3.0 - float(rng.uniform(0, 3))))
    # and two cards poking out beside the wall (must stay visible)
    s.add_node(pc, translation=(7.5, 0, -2.0))
    s.add_node(pc, translation=(-7.5, 1, -2.0))
    return s.finalize()


def mirrored_instances():
    """negative-determinant node transform: facing test flips (mesh.glsl:94-98); one single-sided material"""
    s = Scene.new()
    pos, idx = grid_mesh(8, 8, lambda u, v: (u * 2 - 1, v * 2 - 1, 0.2 * np.sin(u * 6) * np.cos(v * 5)))
    p = s.add_primitive(pos, idx, 0)
    s.add_node(p, translation=(-1.2, 0, 0))
    s.add_node(p, translation=(1.2, 0, 0), scale=(-1, 1, 1))
    s.add_node(p, translation=(0, 1.5, -1), rotation=(0, 0.38268343, 0, 0.92387953), scale=(0.5, -0.7, 1.3))
    return s.finalize()


def random_soup(n_tris=400, seed=3, spread=3.0, size=0.8, materials=1):
    """random intersecting triangles at random depths (some behind the camera), double sided; `materials` > 1 splits them
    over that many primitives, each with its own random albedo (resolve-pass tests)"""
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-spread, spread, (n_tris, 1, 3)).astype(np.float32)
    centers[:, :, 2] = rng.uniform(-12, 4, (n_tris, 1))
    pos = (centers + rng.uniform(-size, size, (n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
    s = Scene.new()
    if materials <= 1:
        m = s.add_material(double_sided=True)
        p = s.add_primitive(pos, np.arange(n_tris * 3, dtype=np.uint32), m)
        s.add_node(p)
        return s.finalize()
    per = (n_tris + materials - 1) // materials
    for k in range(materials):
        lo, hi = k * per, min(n_tris, (k + 1) * per)
        if lo >= hi:
            break
        albedo = tuple(float(x) for x in rng.uniform(0.0, 1.0, 4))
        if k == 0:
            albedo = (0.0, 1.0, 0.002, 1.0)  # both branches of the transfer function and the exact end points
        m = s.add_material(albedo=albedo, double_sided=True)
        p = s.add_primitive(pos[lo * 3:hi * 3], np.arange((hi - lo) * 3, dtype=np.uint32), m)
        s.add_node(p)
    return s.finalize()


def cull_stress(seed=11, instances=61):
    """many instances of a multi-meshlet soup under random node transforms, with the meshlet AABBs OVERWRITTEN by a mix of
    ordinary, degenerate and non-finite boxes (the cull reads nothing else of a meshlet): zero / negative / huge / tiny
    extents, boxes that straddle or sit on the camera plane, infinities and NaNs.  Exercises every branch of the projection
    (packed shared-reciprocal path, plain-division path, NaN ordering) and the odd tail of the two-draws-per-thread split."""
    rng = np.random.default_rng(seed)
    n_tris = 2600
    centers = rng.uniform(-1.0, 1.0, (n_tris, 1, 3)).astype(np.float32)
    pos = (centers + rng.uniform(-0.15, 0.15, (n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
    s = Scene.new()
    m = s.add_material(double_sided=True)
    p = s.add_primitive(pos, np.arange(n_tris * 3, dtype=np.uint32), m)
    wall, widx = grid_mesh(12, 12, lambda u, v: ((u * 2 - 1) * 5, (v * 2 - 1) * 4, np.zeros_like(u)))
    s.add_node(s.add_primitive(wall, widx, m), translation=(0, 0, 1.0))
    tri = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0]], np.float32)
    s.add_node(s.add_primitive(tri, np.arange(3, dtype=np.uint32), m), translation=(0, 2.0, 2.0))  # one more draw: odd total
    for k in range(instances):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        sc = rng.uniform(0.2, 1.5, 3) * rng.choice([1.0, 1.0, 1.0, -1.0], 3)
        t = rng.uniform(-6, 6, 3)
        t[2] = rng.uniform(-14, 5)
        s.add_node(p, translation=tuple(map(float, t)), rotation=tuple(map(float, q)), scale=tuple(map(float, sc)))
    s.finalize()
    ml = s.primitive(0)["meshlets"]
    n = ml.shape[0]
    ext, cen = ml["aabbExtents"], ml["aabbCenter"]
    special = np.array([0.0, -0.0, 1e-30, 1e-42, 1e20, 3e38, np.inf, -np.inf, np.nan, -1.0, 1.0, 0.5, 2.0**-64, 2.0**63], np.float32)
    for i in range(n):
        kind = i % 7
        if kind == 0:
            continue  # the builder's own box
        if kind == 1:   # degenerate extents
            ext[i] = rng.choice([0.0, 0.0, 1e-30, 0.25], 3)
        elif kind == 2:  # arbitrary specials in random slots
            for a in range(3):
                if rng.random() < 0.5:
                    ext[i][a] = rng.choice(special)
                if rng.random() < 0.3:
                    cen[i][a] = rng.choice(special)
        elif kind == 3:  # big boxes that straddle the camera plane
            ext[i] = rng.uniform(2, 30, 3)
        elif kind == 4:  # exact lattice values: zero clip coordinates and exact quotients become likely
            cen[i] = rng.integers(-2, 3, 3)
            ext[i] = rng.choice([0.5, 1.0, 2.0], 3)
        elif kind == 5:  # thin slabs
            ext[i][rng.integers(0, 3)] = 0.0
        else:            # tiny boxes: high mips never selected, level-0 footprints
            ext[i] = rng.uniform(1e-4, 1e-2, 3)
    return s


def camera(W, H, eye=(0, 0, 3), center=(0, 0, 0)):
    return Camera(W, H).look_at(eye, center)


def icosphere_soup(frequency):
    """(positions, flat u32 index list) of the procedural icosphere, taken back out of its meshlets (copies: the scene may go away)"""
    sphere = Scene.icosphere(frequency)
    src = sphere.primitive(0)
    soup = np.concatenate([src["vertex_indices"][int(m["vertexOffset"]) + src["triangles"][int(m["triangleOffset"]): int(m["triangleOffset"]) + 3 * int(m["triangleCount"])].astype(np.int64)]
                           for m in src["meshlets"]]).astype(np.uint32)
    return src["vertices"]["position"].copy(), soup
