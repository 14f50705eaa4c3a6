"""Seeded synthetic slab-electrode systems in OpenMM's CUDA array layouts.

The geometry and the image-particle recipe follow the reference's only workload,
/root/reference/example/nacl_1m_spce/nacl_constv.py:60-95 (one massless image per real atom at
(x, y, -z) with charge -q; neutral wall atoms get their image offset by +0.001 nm, :82-85), and its
composition (nacl_1m_eq.pdb: 38 Na+, 38 Cl-, 2046 SPC/E waters, 1280 massless neutral wall carbons
out of 7494 real atoms, walls last; box 3.9352 x 4.2600 nm, zmax = 4.2183 nm, nacl_constv.py:28).
Forces are *injected*: the integrator step path (BASELINE.json north_star) starts from OpenMM's
fixed-point force buffer, so a slab carries int64 forces scaled by 2^32
(constVLangevin.cu:10) instead of a force field.

Array layouts, P = num_atoms rounded up to a multiple of 32:
  posq   (P,4) real   x, y, z, charge
  corr   (P,4) real   low bits of the position, "mixed" precision only
  velm   (P,4) mixed  vx, vy, vz, 1/mass  (0 for massless)
  force  (3,P) int64  x-plane, y-plane, z-plane, value = llround(f * 2^32)
Slot order is the original order (identity atomIndex): reals [0,R), cell-i images [i*R,(i+1)*R).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

PRECISIONS = ("single", "mixed", "double")
REAL = {"single": np.float32, "mixed": np.float32, "double": np.float64}
MIXED = {"single": np.float32, "mixed": np.float64, "double": np.float64}

BOLTZ = 8.31446261815324e-3  # kJ/mol/K (OpenMM >= 7.4)

# example/nacl_1m_spce geometry
EXAMPLE_LX, EXAMPLE_LY, EXAMPLE_ZMAX = 3.9352, 4.2600, 4.2183
EXAMPLE_REAL_ATOMS = 7494
WALL_FRACTION = 1280.0 / 7494.0
ION_FRACTION = 76.0 / 7494.0

MASS_O, MASS_H, MASS_NA, MASS_CL = 15.9994, 1.008, 22.98977, 35.45
Q_O, Q_H = -0.8476, 0.4238
FORCE_SCALE = 4294967296.0


@dataclass
class Slab:
    precision: str
    num_cells: int
    num_real: int
    num_atoms: int
    padded: int
    zmax: float
    box: tuple
    posq: np.ndarray
    corr: np.ndarray | None
    velm: np.ndarray
    force: np.ndarray
    masses: np.ndarray  # (num_atoms,) float64, original order
    global_atom_offset: int = 0
    meta: dict = field(default_factory=dict)

    def copy(self) -> "Slab":
        return Slab(self.precision, self.num_cells, self.num_real, self.num_atoms, self.padded,
                    self.zmax, self.box, self.posq.copy(),
                    None if self.corr is None else self.corr.copy(), self.velm.copy(),
                    self.force.copy(), self.masses.copy(), self.global_atom_offset,
                    dict(self.meta))

    # -- pair classes, for the algorithmic byte count (DESIGN.md "bytes per pair") --------------
    def pair_classes(self) -> dict:
        w = self.velm[: self.num_real, 3] != 0
        q = self.posq[: self.num_real, 3] != 0
        return {
            "massive_charged": int(np.sum(w & q)),
            "massive_neutral": int(np.sum(w & ~q)),
            "massless_charged": int(np.sum(~w & q)),
            "massless_neutral": int(np.sum(~w & ~q)),
        }


def padded_atoms(n: int) -> int:
    return (n + 31) // 32 * 32


def make_slab(num_real: int, precision: str = "single", num_cells: int = 2, seed: int = 12345,
              temperature: float = 300.0, wall_fraction: float = WALL_FRACTION,
              ion_fraction: float = ION_FRACTION, force_sigma: float = 1000.0,
              image_force_garbage: bool = True, zmax: float = EXAMPLE_ZMAX, x_offset: float = 0.0) -> Slab:
    """Build a seeded slab with ``num_real`` real atoms and ``num_cells - 1`` image cells.
    ``x_offset`` shifts the slab along x (a lateral tile of a wider electrode, see ``slab_tile``)."""
    assert precision in PRECISIONS
    assert num_cells >= 2 and num_cells % 2 == 0
    rng = np.random.Generator(np.random.Philox(seed))
    R = int(num_real)
    N = R * num_cells
    P = padded_atoms(N)
    real, mixed = REAL[precision], MIXED[precision]

    n_wall = int(round(R * wall_fraction))
    n_ion = int(round(R * ion_fraction))
    n_water = (R - n_wall - n_ion) // 3 * 3
    n_ion = R - n_wall - n_water  # absorb the remainder so the counts add up
    scale = np.sqrt(max(R, 1) / EXAMPLE_REAL_ATOMS)  # constant number density in x, y
    lx, ly = EXAMPLE_LX * scale, EXAMPLE_LY * scale

    mass = np.zeros(R, np.float64)
    charge = np.zeros(R, np.float64)
    ion_sign = np.where(np.arange(n_ion) % 2 == 0, 1.0, -1.0)
    mass[:n_ion] = np.where(ion_sign > 0, MASS_NA, MASS_CL)
    charge[:n_ion] = ion_sign
    wsl = slice(n_ion, n_ion + n_water)
    mass[wsl] = np.tile([MASS_O, MASS_H, MASS_H], n_water // 3)
    charge[wsl] = np.tile([Q_O, Q_H, Q_H], n_water // 3)
    # walls: mass 0, charge 0 (spce.xml:28,42-44), split between z = 0 and z = zmax

    pos = np.empty((R, 3), np.float64)
    pos[:, 0] = x_offset + rng.random(R) * lx
    pos[:, 1] = rng.random(R) * ly
    pos[:, 2] = (0.05 + 0.9 * rng.random(R)) * zmax
    wall = slice(n_ion + n_water, R)
    pos[wall, 2] = np.where(np.arange(n_wall) % 2 == 0, 0.0, zmax)

    inv_mass = np.zeros(R, np.float64)
    massive = mass > 0
    inv_mass[massive] = 1.0 / mass[massive]
    vel = rng.standard_normal((R, 3)) * np.sqrt(BOLTZ * temperature * inv_mass)[:, None]

    posq64 = np.zeros((P, 4), np.float64)
    posq64[:R, :3] = pos
    posq64[:R, 3] = charge
    for i in range(1, num_cells):
        img = slice(i * R, (i + 1) * R)
        # initial image placement as in nacl_constv.py:75-87: cell 1 at -z; further cells follow
        # the kernel's recurrence.  The first step rewrites every charged image anyway.
        zi = pos[:, 2].copy()
        for j in range(1, i + 1):
            zi = -zi + 2.0 * j * zmax if j > 1 else -zi
        posq64[img, 0] = pos[:, 0]
        posq64[img, 1] = pos[:, 1]
        posq64[img, 2] = np.where(charge != 0, zi, zi + 0.001)
        posq64[img, 3] = -charge if i % 2 == 1 else charge

    posq = posq64.astype(real)
    corr = None
    if precision == "mixed":
        corr = np.zeros((P, 4), real)
        corr[:, :3] = (posq64[:, :3] - posq[:, :3].astype(np.float64)).astype(real)

    velm = np.zeros((P, 4), mixed)
    velm[:R, :3] = vel.astype(mixed)
    velm[:R, 3] = inv_mass.astype(mixed)

    force = np.zeros((3, P), np.int64)
    f = rng.standard_normal((3, R)) * force_sigma
    force[:, :R] = np.rint(f * FORCE_SCALE).astype(np.int64)
    if image_force_garbage:
        g = rng.standard_normal((3, N - R)) * force_sigma
        force[:, R:N] = np.rint(g * FORCE_SCALE).astype(np.int64)

    masses = np.zeros(N, np.float64)
    masses[:R] = mass
    return Slab(precision, num_cells, R, N, P, float(zmax), (lx, ly, num_cells * zmax), posq, corr,
                velm, force, masses, 0,
                {"seed": seed, "n_ion": n_ion, "n_water": n_water, "n_wall": n_wall,
                 "temperature": temperature})


MASS_DRUDE = 0.4  # amu, the usual Drude particle mass


def make_drude_slab(num_real: int, precision: str = "single", num_cells: int = 2, seed: int = 12345,
                    drude_sigma: float = 0.012, **kw):
    """A slab for ConstVDrudeLangevinIntegrator: ``make_slab`` whose "water" atoms are regrouped as
    polarizable 4-site molecules (parent O carrying a Drude particle on a spring, two H).

    Returns ``(slab, normal_particles, pair_particles)`` as the reference's kernel builds them from the
    DrudeForce (CudaConstVKernels.cpp:213-230): ``pair_particles`` is an (n, 2) int32 array of
    (Drude particle, parent) ORIGINAL indices, ``normal_particles`` the sorted indices of every other
    real atom (ions, hydrogens, massless walls).  The Drude particle sits ``drude_sigma`` nm (rms per
    component) from its parent, so with a hard wall at ~0.02 nm a good fraction of the pairs bounce.
    """
    s = make_slab(num_real, precision, num_cells=num_cells, seed=seed, **kw)
    rng = np.random.Generator(np.random.Philox(seed + 99))
    R = s.num_real
    n_ion, n_water = s.meta["n_ion"], s.meta["n_water"]
    n_mol = n_water // 4
    first = n_ion
    real, mixed = REAL[precision], MIXED[precision]
    parents = first + 4 * np.arange(n_mol)
    drudes = parents + 1
    mass = s.masses[:R].copy()
    charge = s.posq[:R, 3].astype(np.float64)
    pos = s.posq[:R, :3].astype(np.float64)
    if s.corr is not None:
        pos += s.corr[:R, :3].astype(np.float64)
    for k, (m, q) in enumerate(((MASS_O - MASS_DRUDE, 1.0), (MASS_DRUDE, -1.8476), (MASS_H, Q_H), (MASS_H, Q_H))):
        mass[first + k: first + 4 * n_mol: 4] = m
        charge[first + k: first + 4 * n_mol: 4] = q
    # atoms of the former 3-site waters left over after regrouping become neutral, massless fillers
    mass[first + 4 * n_mol: first + n_water] = 0.0
    charge[first + 4 * n_mol: first + n_water] = 0.0
    pos[drudes] = pos[parents] + rng.standard_normal((n_mol, 3)) * drude_sigma
    inv_mass = np.where(mass > 0, 1.0 / np.where(mass > 0, mass, 1.0), 0.0)
    vel = rng.standard_normal((R, 3)) * np.sqrt(BOLTZ * s.meta["temperature"] * inv_mass)[:, None]
    s.masses[:R] = mass
    p64 = np.zeros((s.padded, 4), np.float64)
    p64[:, :3] = s.posq[:, :3]
    if s.corr is not None:
        p64[:, :3] += s.corr[:, :3]
    p64[:, 3] = s.posq[:, 3]
    p64[:R, :3] = pos
    p64[:R, 3] = charge
    for i in range(1, num_cells):
        img = slice(i * R, (i + 1) * R)
        zi = pos[:, 2].copy()
        for j in range(1, i + 1):
            zi = -zi + 2.0 * j * s.zmax if j > 1 else -zi
        p64[img, 0], p64[img, 1] = pos[:, 0], pos[:, 1]
        p64[img, 2] = np.where(charge != 0, zi, zi + 0.001)
        p64[img, 3] = -charge if i % 2 == 1 else charge
    s.posq = p64.astype(real)
    if s.corr is not None:
        s.corr[:, :3] = (p64[:, :3] - s.posq[:, :3].astype(np.float64)).astype(real)
    s.velm[:R, :3] = vel.astype(mixed)
    s.velm[:R, 3] = inv_mass.astype(mixed)
    pairs = np.stack([drudes, parents], axis=1).astype(np.int32)
    in_pair = np.zeros(R, bool)
    in_pair[pairs.reshape(-1)] = True
    normal = np.nonzero(~in_pair)[0].astype(np.int32)
    s.meta.update(n_pairs=int(n_mol), n_normal=int(len(normal)))
    return s, normal, np.ascontiguousarray(pairs)


D_OH = 0.1                                                   # nm, SPC/E (spce.xml)
ANGLE_HOH = np.deg2rad(109.47)
D_HH = float(2.0 * D_OH * np.sin(0.5 * ANGLE_HOH))           # 0.16330 nm


def make_water_slab(num_real: int, precision: str = "single", num_cells: int = 2, seed: int = 12345, **kw):
    """``make_slab`` whose water atoms form RIGID SPC/E molecules (O-H 0.1 nm, H-O-H 109.47 degrees) in
    random orientations: the system of the reference's example, ``constraints=AllBonds`` on 3-site water
    (example/nacl_1m_spce/nacl_constv.py:56), whose constraint hand-off is SETTLE.

    Returns ``(slab, cluster_atoms, cluster_params, constraints)``: ``cluster_atoms`` (n, 3) int32 =
    (O, H1, H2) original indices, ``cluster_params`` (n, 2) float32 = (d_OH, d_HH), and ``constraints`` =
    (p1, p2, distance) arrays as System.getConstraintParameters would list them (O-H1, O-H2, H1-H2).
    """
    s = make_slab(num_real, precision, num_cells=num_cells, seed=seed, **kw)
    rng = np.random.Generator(np.random.Philox(seed + 77))
    R = s.num_real
    n_ion, n_water = s.meta["n_ion"], s.meta["n_water"]
    n_mol = n_water // 3
    real = REAL[precision]
    oxy = n_ion + 3 * np.arange(n_mol)
    p64 = s.posq[:, :3].astype(np.float64)
    if s.corr is not None:
        p64 += s.corr[:, :3]
    u = rng.standard_normal((n_mol, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    t = rng.standard_normal((n_mol, 3))
    w = t - np.einsum("nk,nk->n", t, u)[:, None] * u
    w /= np.linalg.norm(w, axis=1)[:, None]
    c, sn = np.cos(0.5 * ANGLE_HOH), np.sin(0.5 * ANGLE_HOH)
    p64[oxy + 1] = p64[oxy] + D_OH * (c * u + sn * w)
    p64[oxy + 2] = p64[oxy] + D_OH * (c * u - sn * w)
    charge = s.posq[:R, 3].astype(np.float64)
    for i in range(1, num_cells):
        img = slice(i * R, (i + 1) * R)
        zi = p64[:R, 2].copy()
        for j in range(1, i + 1):
            zi = -zi + 2.0 * j * s.zmax if j > 1 else -zi
        p64[img, 0], p64[img, 1] = p64[:R, 0], p64[:R, 1]
        p64[img, 2] = np.where(charge != 0, zi, zi + 0.001)
    s.posq[:, :3] = p64.astype(real)
    if s.corr is not None:
        s.corr[:, :3] = (p64 - s.posq[:, :3].astype(np.float64)).astype(real)
    atoms = np.stack([oxy, oxy + 1, oxy + 2], axis=1).astype(np.int32)
    params = np.tile(np.asarray([D_OH, D_HH], np.float32), (n_mol, 1))
    p1 = np.concatenate([oxy, oxy, oxy + 1]).astype(np.int32)
    p2 = np.concatenate([oxy + 1, oxy + 2, oxy + 2]).astype(np.int32)
    dist = np.concatenate([np.full(n_mol, float(np.float32(D_OH))), np.full(n_mol, float(np.float32(D_OH))),
                           np.full(n_mol, float(np.float32(D_HH)))])
    s.meta.update(n_molecules=int(n_mol))
    return s, np.ascontiguousarray(atoms), np.ascontiguousarray(params), (p1, p2, dist)


def example_system(path: str, precision: str = "mixed", num_cells: int = 2):
    """The reference's example workload, example/nacl_1m_spce (BASELINE.json configs[0-1]): 38 Na+, 38 Cl-,
    2046 SPC/E waters and 1280 massless neutral wall carbons between two electrodes, from the fixture
    tests/golden/nacl_1m_spce_system.npz (the real atoms as `forcefield.createSystem(topology,
    constraints=AllBonds)` sees them; written by tests/golden/make_example_system.py from the PDB and spce.xml).

    The image particles are added here exactly as the example's script adds them (nacl_constv.py:60-95): one
    per real atom, appended after all real atoms, mass 0, charge -q, at (x, y, -z) -- a neutral atom's image at
    (x, y, -z + 0.001 nm) so that it does not sit on its partner (:82-85).  Velocities start at zero (the
    script never sets them); forces are left zero for the caller to inject.  Default precision is the
    script's ('CudaPrecision': 'mixed', :150).

    Returns ``(slab, cluster_atoms, cluster_params, constraints)`` like ``make_water_slab``."""
    assert precision in PRECISIONS and num_cells == 2
    d = np.load(path)
    pos = d["pos_mA"].astype(np.float64) / 1.0e4  # thousandths of an Angstrom -> nm
    mass, charge = d["mass"].astype(np.float64), d["charge"].astype(np.float64)
    R = len(mass)
    N, P = 2 * R, padded_atoms(2 * R)
    real, mixed = REAL[precision], MIXED[precision]
    zmax = float(d["zmax_nm"])
    posq64 = np.zeros((P, 4), np.float64)
    posq64[:R, :3], posq64[:R, 3] = pos, charge
    posq64[R:N, 0], posq64[R:N, 1] = pos[:, 0], pos[:, 1]
    posq64[R:N, 2] = np.where(charge != -charge, -pos[:, 2], -pos[:, 2] + 0.001)
    posq64[R:N, 3] = -charge
    posq = posq64.astype(real)
    corr = None
    if precision == "mixed":
        corr = np.zeros((P, 4), real)
        corr[:, :3] = (posq64[:, :3] - posq[:, :3].astype(np.float64)).astype(real)
    velm = np.zeros((P, 4), mixed)
    velm[:R, 3] = np.where(mass > 0, 1.0 / np.where(mass > 0, mass, 1.0), 0.0).astype(mixed)
    masses = np.zeros(N, np.float64)
    masses[:R] = mass
    p1, p2, dist = d["constraint_p1"], d["constraint_p2"], d["constraint_distance"]
    # the rigid molecules: apex = the atom with two bonds to the others (O), as constv_settle_find_clusters finds them
    n_mol = len(p1) // 3
    oxy = p1[:n_mol * 2:2]
    assert np.array_equal(p1[1:n_mol * 2:2], oxy)
    atoms = np.stack([oxy, p2[:n_mol * 2:2], p2[1:n_mol * 2:2]], axis=1).astype(np.int32)
    assert np.array_equal(p1[2 * n_mol:], atoms[:, 1]) and np.array_equal(p2[2 * n_mol:], atoms[:, 2])
    params = np.stack([dist[:n_mol * 2:2], dist[2 * n_mol:]], axis=1).astype(np.float32)
    box = tuple(float(x) for x in d["box_nm"])
    slab = Slab(precision, 2, R, N, P, zmax, box, posq, corr, velm, np.zeros((3, P), np.int64), masses, 0,
                {"seed": 0, "n_ion": int(np.sum(np.abs(charge) == 1.0)), "n_water": 3 * n_mol,
                 "n_wall": int(np.sum(mass == 0)), "n_molecules": n_mol, "temperature": 300.0,
                 "source": "example/nacl_1m_spce"})
    return slab, np.ascontiguousarray(atoms), np.ascontiguousarray(params), (p1.copy(), p2.copy(), dist.copy())


def permute_slab(slab: Slab, perm) -> tuple:
    """A copy of ``slab`` whose slot ``k`` holds original atom ``perm[k]`` (what OpenMM's ``reorderAtoms``
    leaves behind), plus the padded ``atom_index`` array (padding slots stay in place)."""
    t = slab.copy()
    n = slab.num_atoms
    perm = np.asarray(perm)
    t.posq[:n] = slab.posq[perm]
    if slab.corr is not None:
        t.corr[:n] = slab.corr[perm]
    t.velm[:n] = slab.velm[perm]
    t.force[:, :n] = slab.force[:, perm]
    atom_index = np.arange(slab.padded, dtype=np.int32)
    atom_index[:n] = perm
    return t, atom_index


def molecule_permutation(slab: Slab, first: int, atoms_per_molecule: int, num_molecules: int, window: int = 1024,
                         seed: int = 1) -> np.ndarray:
    """A slot permutation of the kind OpenMM produces: whole identical molecules trade places, real ones among
    the real slots and their images among the image slots of the same cell, and -- because the re-sort is
    spatial -- only with molecules nearby: molecules are shuffled inside windows of ``window`` molecules.
    ``first`` = slot of the first molecule; other atoms (ions, walls) stay where they are."""
    rng = np.random.default_rng(seed)
    R, N = slab.num_real, slab.num_atoms
    perm = np.arange(N, dtype=np.int32)
    for cell in range(slab.num_cells):
        order = np.arange(num_molecules)
        for lo in range(0, num_molecules, window):
            hi = min(num_molecules, lo + window)
            order[lo:hi] = lo + rng.permutation(hi - lo)
        for k in range(atoms_per_molecule):
            dst = cell * R + first + k + atoms_per_molecule * np.arange(num_molecules)
            perm[dst] = cell * R + first + k + atoms_per_molecule * order
    return perm


def slice_slab(slab: Slab, lo: int, hi: int) -> Slab:
    """The shard holding real atoms [lo, hi) and their images (multi-GPU range partition:
    every image lives with its real partner)."""
    R, nc = slab.num_real, slab.num_cells
    r = hi - lo
    n = r * nc
    P = padded_atoms(n)
    src = np.concatenate([np.arange(lo, hi) + i * R for i in range(nc)])
    posq = np.zeros((P, 4), slab.posq.dtype)
    posq[:n] = slab.posq[src]
    corr = None
    if slab.corr is not None:
        corr = np.zeros((P, 4), slab.corr.dtype)
        corr[:n] = slab.corr[src]
    velm = np.zeros((P, 4), slab.velm.dtype)
    velm[:n] = slab.velm[src]
    force = np.zeros((3, P), np.int64)
    force[:, :n] = slab.force[:, src]
    return Slab(slab.precision, nc, r, n, P, slab.zmax, slab.box, posq, corr, velm, force,
                slab.masses[src].copy(), slab.global_atom_offset + lo, dict(slab.meta))


def tile_real_atoms(num_real_total: int, tiles: int) -> int:
    """Real atoms per lateral tile of a tiled slab (``slab_tile``); the total must divide evenly and a tile must
    be a multiple of 32 atoms so that tile boundaries are partition boundaries (``partition.shard_range``)."""
    if tiles < 1 or num_real_total % tiles != 0 or (tiles > 1 and (num_real_total // tiles) % 32 != 0):
        raise ValueError(f"{num_real_total} real atoms do not cut into {tiles} tiles of a multiple of 32 atoms")
    return num_real_total // tiles


def slab_tile(num_real_total: int, tiles: int, t: int, precision: str = "single", seed: int = 12345,
              maker=None, **kw):
    """Lateral tile ``t`` of the seeded global slab of ``num_real_total`` real atoms made of ``tiles`` tiles.

    The multi-GPU workloads are ONE global slab: an electrode ``tiles`` times as wide in x as the example's
    aspect, tile ``t`` occupying x in [t*lx, (t+1)*lx) with its own ions, water and wall atoms (seed + t), so
    that every contiguous real-atom range of one tile has the composition of the whole.  Its real atoms are
    numbered [t*Rt, (t+1)*Rt) in the global slab (``global_atom_offset``), which is exactly
    ``partition.shard_range(num_real_total, t, tiles)``: a rank of a ``tiles``-GPU job builds only its own
    tile, and ``concat_slabs`` of all tiles is the unpartitioned slab (``slice_slab`` of which is the tile).
    ``maker`` = ``make_slab`` (default), ``make_water_slab`` or ``make_drude_slab``; returns what it returns.
    """
    rt = tile_real_atoms(num_real_total, tiles)
    maker = maker or make_slab
    scale = np.sqrt(max(rt, 1) / EXAMPLE_REAL_ATOMS)
    out = maker(rt, precision, seed=seed + t, x_offset=t * EXAMPLE_LX * scale, **kw)
    s = out[0] if isinstance(out, tuple) else out
    s.global_atom_offset = t * rt
    s.meta["tile"] = (t, tiles)
    return out


def concat_slabs(tiles: list) -> Slab:
    """The unpartitioned slab whose real atoms are the tiles' real atoms in order (and likewise each image
    cell): the inverse of ``slice_slab`` over consecutive ranges."""
    first = tiles[0]
    nc = first.num_cells
    R = sum(t.num_real for t in tiles)
    N = R * nc
    P = padded_atoms(N)
    posq = np.zeros((P, 4), first.posq.dtype)
    corr = None if first.corr is None else np.zeros((P, 4), first.corr.dtype)
    velm = np.zeros((P, 4), first.velm.dtype)
    force = np.zeros((3, P), np.int64)
    masses = np.zeros(N, np.float64)
    lo = 0
    for t in tiles:
        r = t.num_real
        for cell in range(nc):
            dst = slice(cell * R + lo, cell * R + lo + r)
            src = slice(cell * r, (cell + 1) * r)
            posq[dst] = t.posq[src]
            if corr is not None:
                corr[dst] = t.corr[src]
            velm[dst] = t.velm[src]
            force[:, dst] = t.force[:, src]
            masses[dst] = t.masses[src]
        lo += r
    box = (sum(t.box[0] for t in tiles), first.box[1], first.box[2])
    return Slab(first.precision, nc, R, N, P, first.zmax, box, posq, corr, velm, force, masses,
                first.global_atom_offset, dict(first.meta))


def new_forces(slab: Slab, step_index: int, force_sigma: float = 1000.0) -> np.ndarray:
    """A fresh (3,P) injected force buffer for a given step (real and image slots non-zero)."""
    rng = np.random.Generator(np.random.Philox(key=slab.meta.get("seed", 0) + 7919 * (step_index + 1)))
    force = np.zeros((3, slab.padded), np.int64)
    f = rng.standard_normal((3, slab.num_atoms)) * force_sigma
    force[:, : slab.num_atoms] = np.rint(f * FORCE_SCALE).astype(np.int64)
    return force


def algorithmic_bytes_per_step(slab: Slab, zero_images: bool = True, rigid_molecules: int = 0,
                               drude_pairs: int = 0, drude: bool = False) -> int:
    """Bytes the fused step must move for this slab (DESIGN.md "bytes per pair").

    ``rigid_molecules`` > 0 (the step that solves the constraints of rigid 3-site molecules itself) adds
    the work-item list, 8 B per ordinary atom or molecule, and the molecule list, 16 + 8 B per molecule.

    single precision, numCells = 2, per real/image pair:
      always read   velm 16 + posq 16
      massive       + read force 24, write velm 16 + posq 16
      charged       + read image charge 4, write image posq 16
      zeroing       + write image velm 16 + image force 24
    => massive & charged pair = 148 B (74 B per atom-step).
    """
    r = np.dtype(REAL[slab.precision]).itemsize
    m = np.dtype(MIXED[slab.precision]).itemsize
    pos_b = 4 * r * (2 if slab.precision == "mixed" else 1)  # posq (+ correction)
    vel_b = 4 * m
    c = slab.pair_classes()
    n_img = slab.num_cells - 1
    massive = c["massive_charged"] + c["massive_neutral"]
    charged = c["massive_charged"] + c["massless_charged"]
    total = slab.num_real * (vel_b + pos_b)
    total += massive * (24 + vel_b + pos_b)
    total += charged * n_img * (r + pos_b)
    if zero_images:
        total += slab.num_real * n_img * (vel_b + 24)
    if drude:  # role table: 4 B per real atom; pair list: 8 B per pair
        total += 4 * slab.num_real + 8 * drude_pairs
    if rigid_molecules:  # work-item list: 8 B per ordinary atom or molecule; molecule list: 16 + 8 B per molecule
        total += 8 * (slab.num_real - 2 * rigid_molecules) + 24 * rigid_molecules
    return int(total)
