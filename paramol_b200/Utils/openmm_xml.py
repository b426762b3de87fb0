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
    for force in root.find("Forces"):
        ftype = force.get("type")
        system.force_names.append(ftype)
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
            system.torsion_idx = np.asarray([[int(b.get("p1")), int(b.get("p2")), int(b.get("p3")), int(b.get("p4"))]
                                             for b in items], dtype=np.int32).reshape(-1, 4)
            system.torsion_per = np.asarray([int(b.get("periodicity")) for b in items], dtype=np.int32)
            system.torsion_par = np.asarray([[float(b.get("phase")), float(b.get("k"))] for b in items]).reshape(-1, 2)
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


def write_system_xml(system, file_name):
    """Serialise an MMSystem in the OpenMM System XML vocabulary (checkpoint hook of ObjectiveFunction.f)."""
    r = repr
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
    out.append('\t\t<Force forceGroup="0" type="HarmonicBondForce" usesPeriodic="0" version="2">\n\t\t\t<Bonds>')
    for (a, b), (d, k) in zip(system.bond_idx, system.bond_par):
        out.append('\t\t\t\t<Bond d="%s" k="%s" p1="%d" p2="%d"/>' % (r(float(d)), r(float(k)), a, b))
    out.append("\t\t\t</Bonds>\n\t\t</Force>")
    out.append('\t\t<Force forceGroup="1" type="HarmonicAngleForce" usesPeriodic="0" version="2">\n\t\t\t<Angles>')
    for (a, b, c), (t, k) in zip(system.angle_idx, system.angle_par):
        out.append('\t\t\t\t<Angle a="%s" k="%s" p1="%d" p2="%d" p3="%d"/>' % (r(float(t)), r(float(k)), a, b, c))
    out.append("\t\t\t</Angles>\n\t\t</Force>")
    out.append('\t\t<Force forceGroup="2" type="PeriodicTorsionForce" usesPeriodic="0" version="2">\n\t\t\t<Torsions>')
    for (a, b, c, d), n, (ph, k) in zip(system.torsion_idx, system.torsion_per, system.torsion_par):
        out.append('\t\t\t\t<Torsion k="%s" p1="%d" p2="%d" p3="%d" p4="%d" periodicity="%d" phase="%s"/>'
                   % (r(float(k)), a, b, c, d, n, r(float(ph))))
    out.append("\t\t\t</Torsions>\n\t\t</Force>")
    method = 0 if system.nonbonded_method == "NoCutoff" else 1
    out.append('\t\t<Force cutoff="%s" forceGroup="11" method="%d" rfDielectric="%s" type="NonbondedForce" version="3">'
               % (r(float(system.cutoff)), method, r(float(system.rf_dielectric))))
    out.append("\t\t\t<Particles>")
    for q, s, e in system.particle_par:
        out.append('\t\t\t\t<Particle eps="%s" q="%s" sig="%s"/>' % (r(float(e)), r(float(q)), r(float(s))))
    out.append("\t\t\t</Particles>\n\t\t\t<Exceptions>")
    for (a, b), (q, s, e) in zip(system.exception_idx, system.exception_par):
        out.append('\t\t\t\t<Exception eps="%s" p1="%d" p2="%d" q="%s" sig="%s"/>' % (r(float(e)), a, b, r(float(q)), r(float(s))))
    out.append("\t\t\t</Exceptions>\n\t\t</Force>")
    out.append("\t</Forces>\n</System>")
    with open(file_name, "w") as f:
        f.write("\n".join(out) + "\n")
    return True
