# -*- coding: utf-8 -*-
"""
Description
-----------
Reader/writer for serialized OpenMM ``System`` XML files (the reference deserialises them with
``XmlSerializer.deserializeSystem`` when ``xml_file`` is given, ``ParaMol/MM_engines/openmm.py:233-237``,
and writes them as checkpoints with ``write_system_xml``, ``openmm.py:341-364``).

Only the forces on the objective-evaluation path are understood: HarmonicBondForce, HarmonicAngleForce,
PeriodicTorsionForce, NonbondedForce (method 0 NoCutoff / 1 CutoffNonPeriodic with reaction field) and
CMMotionRemover (ignored, no energy).
"""
import xml.etree.ElementTree as ET
import numpy as np

from ..MM_engines.mm_system import MMSystem


#: NonbondedForce attributes the OpenMM deserializer reads without defaults (NonbondedForceProxy, version 3), with the values
#: OpenMM itself writes for a NoCutoff / CutoffNonPeriodic force (cf. Tests/aniline.xml:108 in the reference)
_NB_DEFAULTS = (("alpha", "0"), ("cutoff", None), ("dispersionCorrection", "1"), ("ewaldTolerance", ".0005"),
                ("exceptionsUsePeriodic", "0"), ("forceGroup", "0"), ("includeDirectSpace", "1"), ("ljAlpha", "0"), ("ljnx", "0"),
                ("ljny", "0"), ("ljnz", "0"), ("method", None), ("nx", "0"), ("ny", "0"), ("nz", "0"), ("recipForceGroup", "-1"),
                ("rfDielectric", None), ("switchingDistance", "-1"), ("type", "NonbondedForce"), ("useSwitchingFunction", "0"),
                ("version", "3"))


def system_from_xml(file_name):
    root = ET.parse(file_name).getroot()
    particles = root.find("Particles")
    masses = [float(p.get("mass")) for p in particles]
    system = MMSystem(len(masses))
    system.masses = np.asarray(masses)
    box = root.find("PeriodicBoxVectors")
    if box is not None:
        for r, tag in enumerate(("A", "B", "C")):
            v = box.find(tag)
            system.box[r] = [float(v.get("x")), float(v.get("y")), float(v.get("z"))]
    system.force_names = []
    system.force_attrs = []          # the <Force .../> attributes as read, so that a checkpoint writes them back unchanged
    system.torsion_force_sizes = []  # torsions per PeriodicTorsionForce (a System may carry several)
    for force in root.find("Forces"):
        ftype = force.get("type")
        system.force_names.append(ftype)
        system.force_attrs.append(dict(force.attrib))
        if ftype == "HarmonicBondForce":
            items = list(force.find("Bonds"))
            system.bond_idx = np.asarray([[int(b.get("p1")), int(b.get("p2"))] for b in items], dtype=np.int32).reshape(-1, 2)
            system.bond_par = np.asarray([[float(b.get("d")), float(b.get("k"))] for b in items]).reshape(-1, 2)
        elif ftype == "HarmonicAngleForce":
            items = list(force.find("Angles"))
            system.angle_idx = np.asarray([[int(b.get("p1")), int(b.get("p2")), int(b.get("p3"))] for b in items],
                                          dtype=np.int32).reshape(-1, 3)
            system.angle_par = np.asarray([[float(b.get("a")), float(b.get("k"))] for b in items]).reshape(-1, 2)
        elif ftype == "PeriodicTorsionForce":
            items = list(force.find("Torsions"))
            system.add_torsions([[int(b.get("p1")), int(b.get("p2")), int(b.get("p3")), int(b.get("p4"))] for b in items],
                                [int(b.get("periodicity")) for b in items],
                                [[float(b.get("phase")), float(b.get("k"))] for b in items])
            system.torsion_force_sizes.append(len(items))
        elif ftype == "NonbondedForce":
            method = int(force.get("method"))
            if method not in (0, 1):
                raise NotImplementedError("NonbondedForce method %d (periodic) is outside the path" % method)
            system.nonbonded_method = "NoCutoff" if method == 0 else "CutoffNonPeriodic"
            system.cutoff = float(force.get("cutoff"))
            system.rf_dielectric = float(force.get("rfDielectric"))
            items = list(force.find("Particles"))
            system.particle_par = np.asarray([[float(b.get("q")), float(b.get("sig")), float(b.get("eps"))] for b in items]).reshape(-1, 3)
            items = list(force.find("Exceptions"))
            system.exception_idx = np.asarray([[int(b.get("p1")), int(b.get("p2"))] for b in items], dtype=np.int32).reshape(-1, 2)
            system.exception_par = np.asarray([[float(b.get("q")), float(b.get("sig")), float(b.get("eps"))] for b in items]).reshape(-1, 3)
    return system


def _attrs(pairs):
    return " ".join('%s="%s"' % (k, v) for k, v in pairs)


def write_system_xml(system, file_name):
    """
    Serialise an MMSystem as an OpenMM System XML (the checkpoint hook of ``ObjectiveFunction.f``; the reference writes
    ``XmlSerializer.serialize(system)``, ``ParaMol/MM_engines/openmm.py:341-364``).  Forces are written in the order of
    ``system.force_names`` -- every force of the System, including CMMotionRemover and additional PeriodicTorsionForces, none
    invented -- with the attributes they were read with (``system.force_attrs``) or OpenMM's own defaults, and the
    NonbondedForce carries the complete version-3 vocabulary (alpha, dispersionCorrection, ewaldTolerance, ..., the empty
    GlobalParameters / ParticleOffsets / ExceptionOffsets nodes) that ``XmlSerializer.deserialize`` reads without defaults.
    """
    r = repr
    names = list(getattr(system, "force_names", None) or ["HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce",
                                                          "NonbondedForce", "CMMotionRemover"])
    read_attrs = list(getattr(system, "force_attrs", None) or [])
    tors_sizes = list(getattr(system, "torsion_force_sizes", None) or [])
    n_tors_forces = sum(1 for n in names if n == "PeriodicTorsionForce")
    if len(tors_sizes) != n_tors_forces or sum(tors_sizes) != len(system.torsion_idx):
        # torsions added after reading (add_torsions) belong to the last PeriodicTorsionForce
        tors_sizes = [0] * max(n_tors_forces - 1, 0) + [len(system.torsion_idx)] if n_tors_forces else []
        if read_attrs and getattr(system, "torsion_force_sizes", None) and n_tors_forces == len(system.torsion_force_sizes):
            tors_sizes = list(system.torsion_force_sizes)
            tors_sizes[-1] += len(system.torsion_idx) - sum(system.torsion_force_sizes)
    out = ['<?xml version="1.0" ?>', '<System openmmVersion="7.7" type="System" version="1">', "\t<PeriodicBoxVectors>"]
    for tag, v in zip(("A", "B", "C"), system.box):
        out.append('\t\t<%s x="%s" y="%s" z="%s"/>' % (tag, r(float(v[0])), r(float(v[1])), r(float(v[2]))))
    out.append("\t</PeriodicBoxVectors>")
    out.append("\t<Particles>")
    for m in system.masses:
        out.append('\t\t<Particle mass="%s"/>' % r(float(m)))
    out.append("\t</Particles>")
    out.append("\t<Constraints/>")
    out.append("\t<Forces>")
    tors_lo = 0
    tors_seen = 0
    for k, name in enumerate(names):
        given = read_attrs[k] if k < len(read_attrs) and read_attrs[k].get("type") == name else {}
        group = given.get("forceGroup", str(k))
        if name == "HarmonicBondForce":
            out.append('\t\t<Force %s>\n\t\t\t<Bonds>' % _attrs((("forceGroup", group), ("type", name),
                                                                ("usesPeriodic", given.get("usesPeriodic", "0")), ("version", "2"))))
            for (a, b), (d, kk) in zip(system.bond_idx, system.bond_par):
                out.append('\t\t\t\t<Bond d="%s" k="%s" p1="%d" p2="%d"/>' % (r(float(d)), r(float(kk)), a, b))
            out.append("\t\t\t</Bonds>\n\t\t</Force>")
        elif name == "HarmonicAngleForce":
            out.append('\t\t<Force %s>\n\t\t\t<Angles>' % _attrs((("forceGroup", group), ("type", name),
                                                                 ("usesPeriodic", given.get("usesPeriodic", "0")), ("version", "2"))))
            for (a, b, c), (t, kk) in zip(system.angle_idx, system.angle_par):
                out.append('\t\t\t\t<Angle a="%s" k="%s" p1="%d" p2="%d" p3="%d"/>' % (r(float(t)), r(float(kk)), a, b, c))
            out.append("\t\t\t</Angles>\n\t\t</Force>")
        elif name == "PeriodicTorsionForce":
            n_here = tors_sizes[tors_seen] if tors_seen < len(tors_sizes) else 0
            tors_seen += 1
            out.append('\t\t<Force %s>\n\t\t\t<Torsions>' % _attrs((("forceGroup", group), ("type", name),
                                                                   ("usesPeriodic", given.get("usesPeriodic", "0")), ("version", "2"))))
            for q in range(tors_lo, tors_lo + n_here):
                a, b, c, d = system.torsion_idx[q]
                ph, kk = system.torsion_par[q]
                out.append('\t\t\t\t<Torsion k="%s" p1="%d" p2="%d" p3="%d" p4="%d" periodicity="%d" phase="%s"/>'
                           % (r(float(kk)), a, b, c, d, int(system.torsion_per[q]), r(float(ph))))
            tors_lo += n_here
            out.append("\t\t\t</Torsions>\n\t\t</Force>")
        elif name == "NonbondedForce":
            own = {"cutoff": r(float(system.cutoff)), "method": "0" if system.nonbonded_method == "NoCutoff" else "1",
                   "rfDielectric": r(float(system.rf_dielectric)), "forceGroup": group}
            pairs = [(key, own.get(key, given.get(key, default))) for key, default in _NB_DEFAULTS]
            out.append("\t\t<Force %s>" % _attrs(pairs))
            out.append("\t\t\t<GlobalParameters/>\n\t\t\t<ParticleOffsets/>\n\t\t\t<ExceptionOffsets/>")
            out.append("\t\t\t<Particles>")
            for q, s, e in system.particle_par:
                out.append('\t\t\t\t<Particle eps="%s" q="%s" sig="%s"/>' % (r(float(e)), r(float(q)), r(float(s))))
            out.append("\t\t\t</Particles>\n\t\t\t<Exceptions>")
            for (a, b), (q, s, e) in zip(system.exception_idx, system.exception_par):
                out.append('\t\t\t\t<Exception eps="%s" p1="%d" p2="%d" q="%s" sig="%s"/>' % (r(float(e)), a, b, r(float(q)), r(float(s))))
            out.append("\t\t\t</Exceptions>\n\t\t</Force>")
        elif name == "CMMotionRemover":
            out.append("\t\t<Force %s/>" % _attrs((("forceGroup", group), ("frequency", given.get("frequency", "1")), ("type", name),
                                                    ("version", "1"))))
        else:
            raise NotImplementedError("write_system_xml: force %s is outside the objective-evaluation path" % name)
    out.append("\t</Forces>\n</System>")
    with open(file_name, "w") as f:
        f.write("\n".join(out) + "\n")
    return True
