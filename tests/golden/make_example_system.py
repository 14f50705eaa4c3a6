"""Derives the reference's only workload, example/nacl_1m_spce (BASELINE.json configs[0-1]), WITHOUT OpenMM:
parses nacl_1m_eq.pdb and spce.xml in plain Python and records what `forcefield.createSystem(topology,
constraints=AllBonds)` (nacl_constv.py:56; rigidWater defaults to True) makes of them, as far as the
integrator is concerned -- per real atom its position, mass and charge, plus the distance constraints -- in
a small fixture.  The PDB itself is not copied; the image particles are NOT stored either: they are built at
load time by the example's own recipe (nacl_constv.py:60-95; openmm_constv_b200.slab.example_system).

    python tests/golden/make_example_system.py  ->  tests/golden/nacl_1m_spce_system.npz

What OpenMM does, restated [recalled where marked]:
  * PDBFile: ATOM/HETATM records, columns 13-16 atom name, 18-20 residue name, 23-26 residue number, 31-54
    x, y, z in Angstrom (-> nm), 77-78 element; CRYST1 a, b, c in Angstrom.  Water residues (HOH) get the
    bonds O-H1 and O-H2 from OpenMM's residue table [recalled]; ions and the wall carbons have no bonds.
  * ForceField: a residue is matched to the template with the same elements and bond graph: HOH -> "SPCE"
    (OW, HW, HW), SOD -> "SOD", CLA -> "CLA", grp -> "grp"; mass from the atom type, charge from the
    template atom (spce.xml <AtomTypes>, <Residues>).
  * constraints=AllBonds: every bond becomes a constraint at the HarmonicBondForce length of its classes
    (OW-HW 0.1 nm).  rigidWater=True: each water's H-H distance is constrained too, at
    sqrt(l1^2 + l2^2 - 2 l1 l2 cos(theta)) with theta the HW-OW-HW HarmonicAngleForce angle [recalled].
"""
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/example/nacl_1m_spce"
ZMAX = 4.2183  # nacl_constv.py:28


def parse_forcefield(path):
    root = ET.parse(path).getroot()
    types = {t.get("name"): {"class": t.get("class"), "element": t.get("element"), "mass": float(t.get("mass"))}
             for t in root.find("AtomTypes")}
    residues = {}
    for res in root.find("Residues"):
        atoms = [(a.get("name"), a.get("type"), float(a.get("charge"))) for a in res.findall("Atom")]
        bonds = [(b.get("atomName1"), b.get("atomName2")) for b in res.findall("Bond")]
        residues[res.get("name")] = {"atoms": atoms, "bonds": bonds}
    bond_len = {}
    for b in root.find("HarmonicBondForce"):
        bond_len[frozenset((b.get("class1"), b.get("class2")))] = float(b.get("length"))
    angles = {}
    for a in root.find("HarmonicAngleForce"):
        angles[(a.get("class1"), a.get("class2"), a.get("class3"))] = float(a.get("angle"))
    return types, residues, bond_len, angles


def parse_pdb(path):
    atoms, box = [], None
    for line in open(path):
        rec = line[:6]
        if rec == "CRYST1":
            box = tuple(float(line[6 + 9 * k:15 + 9 * k]) for k in range(3))
        elif rec in ("ATOM  ", "HETATM"):
            atoms.append({"name": line[12:16].strip(), "res": line[17:20].strip(), "resnum": line[22:26],
                          # thousandths of an Angstrom, exactly as printed
                          "xyz_mA": tuple(int(round(float(line[30 + 8 * k:38 + 8 * k]) * 1000)) for k in range(3)),
                          "element": line[76:78].strip()})
    return atoms, box


def match_template(res_atoms, residues, types):
    """ForceField's matching, for this file's four residue kinds: same multiset of elements."""
    elements = sorted(a["element"].upper() for a in res_atoms)
    hits = [name for name, t in residues.items() if sorted(types[ty]["element"].upper() for _, ty, _ in t["atoms"]) == elements]
    by_name = {"HOH": "SPCE", "SOD": "SOD", "CLA": "CLA", "grp": "grp"}[res_atoms[0]["res"]]
    assert by_name in hits, (res_atoms[0]["res"], hits)
    return by_name


def build(ref_dir=REF):
    types, residues, bond_len, angles = parse_forcefield(os.path.join(ref_dir, "spce.xml"))
    atoms, box = parse_pdb(os.path.join(ref_dir, "nacl_1m_eq.pdb"))
    # group into residues in file order
    groups, last = [], None
    for i, a in enumerate(atoms):
        key = (a["res"], a["resnum"])
        if key != last:
            groups.append([])
            last = key
        groups[-1].append(i)
    type_names = sorted({ty for t in residues.values() for _, ty, _ in t["atoms"]})
    n = len(atoms)
    mass, charge, type_index = np.zeros(n), np.zeros(n), np.zeros(n, np.uint8)
    p1, p2, dist = [], [], []
    hh = []
    for g in groups:
        tname = match_template([atoms[i] for i in g], residues, types)
        tmpl = residues[tname]
        # template atoms in the residue's own order by element (O H H / single atoms): map by position
        order = {}
        if tname == "SPCE":
            names = {"O": "OH2", "H1": "H1", "H2": "H2"}
            for i in g:
                order[names[atoms[i]["name"]]] = i
        else:
            order[tmpl["atoms"][0][0]] = g[0]
        for aname, ty, q in tmpl["atoms"]:
            i = order[aname]
            mass[i] = types[ty]["mass"]
            charge[i] = q
            type_index[i] = type_names.index(ty)
        cls = {aname: types[ty]["class"] for aname, ty, _ in tmpl["atoms"]}
        for a, b in tmpl["bonds"]:  # constraints=AllBonds
            p1.append(order[a]); p2.append(order[b]); dist.append(bond_len[frozenset((cls[a], cls[b]))])
        if tname == "SPCE":  # rigidWater: the H-H distance from the two bond lengths and the angle
            l1 = l2 = bond_len[frozenset(("OW", "HW"))]
            theta = angles[("HW", "OW", "HW")]
            hh.append((order["H1"], order["H2"], math.sqrt(l1 * l1 + l2 * l2 - 2 * l1 * l2 * math.cos(theta))))
    for a, b, d in hh:
        p1.append(a); p2.append(b); dist.append(d)
    return {
        "pos_mA": np.asarray([a["xyz_mA"] for a in atoms], np.int32),
        "mass": mass, "charge": charge, "type_index": type_index, "type_names": np.asarray(type_names),
        "constraint_p1": np.asarray(p1, np.int32), "constraint_p2": np.asarray(p2, np.int32),
        "constraint_distance": np.asarray(dist, np.float64),
        "box_nm": np.asarray([box[0] / 10.0, box[1] / 10.0, 2 * ZMAX]),  # nacl_constv.py:50-52: z is set to 2*zmax
        "zmax_nm": np.asarray(ZMAX),
        "source": np.asarray("example/nacl_1m_spce/nacl_1m_eq.pdb + spce.xml via tests/golden/make_example_system.py"),
    }


if __name__ == "__main__":
    data = build(sys.argv[1] if len(sys.argv) > 1 else REF)
    out = os.path.join(HERE, "nacl_1m_spce_system.npz")
    np.savez_compressed(out, **data)
    print(out, os.path.getsize(out), "bytes;", len(data["mass"]), "real atoms,", len(data["constraint_p1"]), "constraints")
