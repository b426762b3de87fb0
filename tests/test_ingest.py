"""Host-side ingest: prmtop/inpcrd/xml/nc/gaussian readers against the reference's fixtures and goldens."""
import os
import numpy as np
import xml.etree.ElementTree as ET

from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
from paramol_b200.Utils.netcdf_lite import read_nc, read_paramol_data, write_paramol_raw
from paramol_b200.Utils.openmm_xml import system_from_xml, write_system_xml
from paramol_b200.Utils.gaussian_esp import GaussianESP

VEC = "test_objective_function.py::test_objective_function_and_properties::optimizable_parameters_values_to_compare"


def test_prmtop_term_counts_and_order_match_xml(aniline_system, golden_dir):
    s = aniline_system
    # reference counts: Tests/Force_field/test_force_field.py:15-45 (14/21/35 terms, 14 particles, 25 1-4 pairs)
    assert (len(s.bond_idx), len(s.angle_idx), len(s.torsion_idx), s.n_atoms, len(s.exception_idx)) == (14, 21, 35, 14, 60)
    assert int(np.sum((np.abs(s.exception_par[:, 0]) >= 1e-8) | (np.abs(s.exception_par[:, 2]) >= 1e-8))) == 25
    # indexing / exclusion lists bit-exact against the serialized OpenMM system (Tests/aniline.xml:26-190)
    x = system_from_xml(os.path.join(golden_dir, "aniline.xml"))
    assert np.array_equal(s.bond_idx, x.bond_idx)
    assert np.array_equal(s.angle_idx, x.angle_idx)
    assert np.array_equal(s.torsion_idx, x.torsion_idx)
    assert np.array_equal(s.torsion_per, x.torsion_per)
    assert np.array_equal(s.exception_idx, x.exception_idx)
    assert len(s.nonexcluded_pairs()) == 91 - 60


def test_prmtop_parameter_values_bit_exact_with_reference_golden_vector(aniline_system, goldens):
    s = aniline_system
    v = np.asarray(goldens[VEC])
    mine = np.concatenate([s.bond_par.ravel(), s.angle_par.ravel(), s.torsion_par.ravel(), s.particle_par.ravel()])
    # the reference prints 17 significant digits: unit conversions reproduce them exactly
    assert np.array_equal(mine, v[:182])


def test_other_molecules_parse(golden_dir):
    counts = {"norfloxacin": (41, 43, 76, 126), "aspirin": (21, 21, 32, 55), "caffeine": (24, 25, 43, 62), "ethane": (8, 7, 12, 9)}
    for name, (n, nb, na, nt) in counts.items():
        s = system_from_prmtop(os.path.join(golden_dir, name + ".prmtop"))
        assert (s.n_atoms, len(s.bond_idx), len(s.angle_idx), len(s.torsion_idx)) == (n, nb, na, nt)
        x = read_inpcrd(os.path.join(golden_dir, name + ".inpcrd"))
        assert x.shape == (n, 3)
        # every pair is either an exception or a plain pair, never both
        assert len(s.nonexcluded_pairs()) + len(s.excluded_pair_set()) == n * (n - 1) // 2


def test_nc_reader(aniline_data, goldens, tmp_path):
    X, E, F = aniline_data
    assert X.shape == (10, 14, 3) and F.shape == (10, 14, 3) and E.shape == (10,)
    # Tests/System/test_system.py:171-226 (filter_conformations leaves the lowest-energy structure = index 0)
    np.testing.assert_almost_equal(X[0], np.asarray(goldens["test_system.py::test_filter_conformations::coordinates_to_compare"])[0])
    np.testing.assert_almost_equal(F[0], np.asarray(goldens["test_system.py::test_filter_conformations::forces_to_compare"])[0])
    np.testing.assert_almost_equal(E[0], goldens["test_system.py::test_filter_conformations::energies_to_compare"][0])
    assert np.abs(F.sum(axis=1)).max() < 1e-9  # net QM force vanishes
    p = str(tmp_path / "rt.nc")
    write_paramol_raw(p, {"reference_coordinates": X, "reference_energies": E, "reference_forces": F})
    X2, E2, F2 = read_paramol_data(p)
    assert np.array_equal(X, X2) and np.array_equal(E, E2) and np.array_equal(F, F2)


def test_xml_roundtrip(golden_dir, tmp_path):
    s = system_from_xml(os.path.join(golden_dir, "aniline.xml"))
    assert s.nonbonded_method == "CutoffNonPeriodic" and abs(s.rf_dielectric - 78.3) < 1e-12
    p = str(tmp_path / "s.xml")
    write_system_xml(s, p)
    s2 = system_from_xml(p)
    for a in ("bond_idx", "bond_par", "angle_idx", "angle_par", "torsion_idx", "torsion_per", "torsion_par",
              "particle_par", "exception_idx", "exception_par"):
        assert np.array_equal(getattr(s, a), getattr(s2, a)), a
    ET.parse(p)


def test_checkpoint_xml_has_the_vocabulary_openmm_deserializes(golden_dir, tmp_path, aniline_system):
    """The checkpoint written by ObjectiveFunction.f must be loadable by XmlSerializer.deserialize like the reference's own
    (ParaMol/MM_engines/openmm.py:341-364).  OpenMM is not installed here, so the check is structural against the reference
    fixture Tests/aniline.xml, which IS XmlSerializer output: same force list in the same order with the same forceGroups, every
    attribute and child node of every <Force> present (the NonbondedForce proxy reads alpha, ewaldTolerance, dispersionCorrection,
    the offsets nodes ... without defaults), identical term data."""
    ref_path = os.path.join(golden_dir, "aniline.xml")
    s = system_from_xml(ref_path)
    p = str(tmp_path / "ck.xml")
    write_system_xml(s, p)
    ref_forces = list(ET.parse(ref_path).getroot().find("Forces"))
    got_forces = list(ET.parse(p).getroot().find("Forces"))
    assert [f.get("type") for f in got_forces] == [f.get("type") for f in ref_forces]      # CMMotionRemover kept, none invented
    for fr, fg in zip(ref_forces, got_forces):
        assert set(fr.attrib) <= set(fg.attrib), (fr.get("type"), set(fr.attrib) - set(fg.attrib))
        assert fr.get("forceGroup") == fg.get("forceGroup") and fr.get("version") == fg.get("version")
        assert [c.tag for c in fr] == [c.tag for c in fg], fr.get("type")
        for cr, cg in zip(fr, fg):
            assert len(cr) == len(cg)
            for er, eg in zip(cr, cg):
                assert er.tag == eg.tag and set(er.attrib) == set(eg.attrib)
                for key in er.attrib:
                    assert float(er.get(key)) == float(eg.get(key)), (fr.get("type"), key)
        if fr.get("type") == "NonbondedForce":
            for key in fr.attrib:
                if key != "type":
                    assert float(fr.get(key)) == float(fg.get(key)), key
    # a System that was not read from XML (AMBER topology): complete vocabulary with OpenMM's defaults, one force per name
    p2 = str(tmp_path / "ck2.xml")
    write_system_xml(aniline_system, p2)
    nb = [f for f in ET.parse(p2).getroot().find("Forces") if f.get("type") == "NonbondedForce"][0]
    nb_ref = [f for f in ref_forces if f.get("type") == "NonbondedForce"][0]
    assert set(nb_ref.attrib) <= set(nb.attrib) and nb.get("method") == "0"
    assert [c.tag for c in nb] == [c.tag for c in nb_ref]
    # a second PeriodicTorsionForce survives the round trip with its own torsions
    s3 = system_from_xml(ref_path)
    s3.force_names.insert(3, "PeriodicTorsionForce")
    s3.force_attrs.insert(3, {"forceGroup": "7", "type": "PeriodicTorsionForce", "usesPeriodic": "0", "version": "2"})
    s3.torsion_force_sizes.append(2)
    s3.add_torsions([[0, 1, 2, 3], [3, 2, 1, 0]], [2, 3], [[0.0, 1.5], [3.0, 0.5]])
    p3 = str(tmp_path / "ck3.xml")
    write_system_xml(s3, p3)
    s4 = system_from_xml(p3)
    assert s4.torsion_force_sizes == [len(s.torsion_idx), 2] and s4.force_names == s3.force_names
    assert np.array_equal(s4.torsion_idx, s3.torsion_idx) and np.array_equal(s4.torsion_par, s3.torsion_par)


def test_gaussian_reader_excerpt(golden_dir):
    conf, grid, esp, energy = GaussianESP().read_log_files(os.path.join(golden_dir, "aniline_opt_excerpt.log"))
    full = np.load(os.path.join(golden_dir, "aniline_esp.npz"))
    assert conf[0].shape == (14, 3) and grid[0].shape == (200, 3) and esp[0].shape == (200,)
    assert np.array_equal(conf[0], full["conformation"])
    assert np.array_equal(grid[0], full["grid"][:200]) and np.array_equal(esp[0], full["esp"][:200])
    assert full["grid"].shape == (39946, 3)  # Tests/Tasks/test_resp.py:78-79
    assert abs(energy[0] - float(full["energy"])) < 1e-9
