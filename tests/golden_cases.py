"""The models behind tests/golden/*.npz (generated from the reference by tests/golden/make_golden.py)."""
from modelspec import (ELASTIC, J2_STEEL, brick_block, brick_periodic_equaldof, cantilever2d, frame2d, frame2d_diaphragm_equaldof,
                       frame3d, quad_plane, quad_plane_stress_pressure, soil_column_equaldof, with_beam_rho, with_pdelta)

# name -> (spec factory, numberer, soe, displacement scale)
CASES = {
    "brick_j2_plain_csc": (lambda: brick_block(3, 3, 2, mat=J2_STEEL, distort=0.25, seed=1, body=(0.0, 0.0, -0.01)), 0, 0, 2e-3),
    "brick_j2_rcm_csr": (lambda: brick_block(4, 2, 3, mat=J2_STEEL, distort=0.2, seed=2), 1, 1, 2e-3),
    "brick_elastic_rcm_csc": (lambda: brick_block(2, 3, 3, mat=ELASTIC, distort=0.3, seed=3), 1, 0, 2e-3),
    "quad_elastic_rcm_csc": (lambda: quad_plane(10, 4, mat=ELASTIC, distort=0.0, seed=4), 1, 0, 2e-2),
    "quad_j2_plain_csr": (lambda: quad_plane(6, 5, mat=J2_STEEL, lx=6.0, ly=5.0, distort=0.3, seed=5, body=(0.0, -0.02)), 0, 1, 2e-3),
    # 2D RC frame: forceBeamColumn + fibre sections (Steel02, Concrete02); scale per dof (ux, uy, rz)
    "frame2d_fiber_rcm_csc": (lambda: frame2d(2, 3, 2), 1, 0, (0.006, 0.003, 6e-5)),
    "frame2d_fiber_plain_csr": (lambda: frame2d(3, 2, 1, nip=4), 0, 1, (0.008, 0.002, 5e-5)),
    # 3D RC space frame: forceBeamColumn (ForceBeamColumn3d) + FiberSection3d; scale per dof (u, r)
    "frame3d_fiber_rcm_csc": (lambda: frame3d(2, 1, 2, ndiv=1), 1, 0, (0.015, 0.015, 0.002, 1e-4, 1e-4, 1e-4)),
    "frame3d_fiber_plain_csr": (lambda: frame3d(1, 2, 2, ndiv=2, nip=5), 0, 1, (0.005, 0.0075, 0.00075, 5e-5, 2.5e-5, 5e-5)),
    # MP constraints (`equalDOF`, PlainHandler's -4 ids): tied dofs inside one element / across elements / onto a
    # fixed dof / several constrained dofs on one retained dof
    "equaldof_soilcolumn_plain_csc": (lambda: soil_column_equaldof(8), 0, 0, 2e-3),
    "equaldof_soilcolumn_rcm_csr": (lambda: soil_column_equaldof(5, mat=ELASTIC, seed=23), 1, 1, 2e-2),
    "equaldof_brick_rcm_csc": (lambda: brick_periodic_equaldof(3, 3, 2), 1, 0, 2e-3),
    "equaldof_brick_plain_csr": (lambda: brick_periodic_equaldof(2, 3, 3, dofs=(0, 1, 2), seed=24), 0, 1, 2e-3),
    "equaldof_frame2d_rcm_csc": (lambda: frame2d_diaphragm_equaldof(2, 2, 1), 1, 0, (0.006, 0.003, 6e-5)),
    "equaldof_frame2d_plain_csr": (lambda: frame2d_diaphragm_equaldof(3, 2, 2, nip=4), 0, 1, (0.008, 0.002, 5e-5)),
    # FourNodeQuad with the PlaneStress material copy (ElasticIsotropicPlaneStress2D) and surface pressure
    "quad_planestress_pressure_rcm_csc": (lambda: quad_plane_stress_pressure(8, 5, 1, 3.0), 1, 0, 2e-2),
    "quad_planestress_j2_rcm_csr": (lambda: quad_plane_stress_pressure(6, 4, 1, -2.0, mat=J2_STEEL, seed=35), 1, 1, 2e-3),
    "quad_planestrain_pressure_j2_plain_csr": (lambda: quad_plane_stress_pressure(6, 4, 0, -2.0, mat=J2_STEEL, seed=32), 0, 1, 2e-3),
}
NSTEPS = 3


def ele_nd(spec):
    """dofs of one element of the (single-kind) model"""
    return {0: 24, 1: 8, 2: 6, 3: 12}[spec.groups[0].kind]


def _uniform_mass(val, rot=None):
    def f(spec):
        import numpy as np
        m = np.full((spec.nn, spec.ndf), val)
        if rot is not None:
            m[:, spec.ndm:] = rot       # rotational dofs: rz in 2D, rx ry rz in 3D
        return m
    return f


# Newmark (displacement form) with nodal masses: name -> (spec factory, mass(spec), gamma, beta, dt)
TRANSIENT_CASES = {
    "newmark_brick_j2": (lambda: brick_block(3, 3, 4, mat=J2_STEEL, lz=3.0, load=(40.0, 0.0, -5.0)), _uniform_mass(0.05), 0.5, 0.25, 0.02),
    "newmark_frame2d": (lambda: frame2d(2, 2, 2, lateral=30.0), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02),
    "newmark_frame3d": (lambda: frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0)), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02),
}


# the same with `rayleigh alphaM betaK betaKinit betaKcomm` on elements and nodes, and element masses from the
# material density (stdBrick: consistent, FourNodeQuad: lumped): name -> (spec, mass(spec), gamma, beta, dt, rayleigh)
J2_STEEL_RHO = (J2_STEEL[0], J2_STEEL[1] + [2.0e-3])
RAYLEIGH = (0.3, 0.002, 0.001, 0.0015)
RAYLEIGH_CASES = {
    "rayleigh_brick_j2": (lambda: brick_block(3, 3, 4, mat=J2_STEEL_RHO, lz=3.0, load=(40.0, 0.0, -5.0), distort=0.2, seed=6),
                          _uniform_mass(0.05), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_quad_j2": (lambda: quad_plane(6, 4, mat=J2_STEEL_RHO, lx=6.0, ly=4.0, distort=0.2, seed=7),
                         _uniform_mass(0.05), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_frame2d": (lambda: frame2d(2, 2, 2, lateral=30.0), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_frame3d": (lambda: frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0)), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_quad_planestress_j2": (lambda: quad_plane_stress_pressure(6, 4, 1, 2.5, mat=J2_STEEL_RHO, seed=36), _uniform_mass(0.05), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_quad_planestress": (lambda: quad_plane_stress_pressure(6, 4, 1, 2.5, seed=33), _uniform_mass(0.05), 0.5, 0.25, 0.02, RAYLEIGH),
    # with `equalDOF`: the nodal masses / nodal unbalance of every dof on a shared equation, in DOF_Group order
    "rayleigh_soilcolumn_equaldof": (lambda: soil_column_equaldof(6, mat=J2_STEEL_RHO), _uniform_mass(0.05), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_frame2d_equaldof": (lambda: frame2d_diaphragm_equaldof(2, 2, 1, lateral=30.0), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    # `geomTransf PDelta`: the geometric stiffness inside Kt and Kc of Element::getDamp, none inside the initial stiffness
    "rayleigh_frame2d_pdelta": (lambda: with_pdelta(frame2d(2, 2, 2, lateral=30.0, gravity=-150.0)), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    # `forceBeamColumn -mass rho`: the elements' lumped mass in the tangent, the inertia and the alphaM damping forces
    "rayleigh_frame2d_rho": (lambda: with_beam_rho(frame2d(2, 2, 2, lateral=30.0), 2.0e-3), _uniform_mass(0.03, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_frame3d_rho": (lambda: with_beam_rho(with_pdelta(frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0), gravity=-60.0)), 1.5e-3), _uniform_mass(0.03, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
    "rayleigh_frame3d_pdelta": (lambda: with_pdelta(frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0), gravity=-90.0)), _uniform_mass(0.05, 0.0), 0.5, 0.25, 0.02, RAYLEIGH),
}


def newmark_coeffs(gamma, beta, dt):
    """c1 c2 c3 and the predictor's a1..a4 of Newmark::newStep (displacement unknown)"""
    return ((1.0, gamma / (beta * dt), 1.0 / (beta * dt * dt)),
            (1.0 - gamma / beta, dt * (1.0 - 0.5 * gamma / beta), -1.0 / (beta * dt), 1.0 - 0.5 / beta))


# integrator DisplacementControl + Newton + NormDispIncr, run by the reference's own classes:
# name -> (spec factory, numberer, soe, control node tag (None: last node), control dof, increment, steps, tol, max_iter)
DISPCONTROL_CASES = {
    # BASELINE configs[0]: the Ex2b cantilever pushover (to 5 % drift), with the RC fibre section
    "dc_cantilever_fiber": (lambda: cantilever2d(ndiv=1), 0, 0, None, 0, 0.432, 50, 1e-8, 10),
    # BASELINE configs[2] in small: J2 brick column pushed under displacement control
    # 3D RC space frame pushed at a roof corner under displacement control (biaxial bending + torsion)
    "dc_frame3d": (lambda: frame3d(1, 1, 2, ndiv=1, lateral=(1.0, 0.6), gravity=-2.0), 1, 0, "roof", 0, 0.25, 9, 3e-7, 12),
    "dc_brick_j2": (lambda: brick_block(3, 3, 5, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.0, 0.0, -0.3)), 1, 0, None, 0, 5e-3, 12, 3e-10, 15),
}


def control_node(spec, node):
    """control node of a DISPCONTROL case: None = the last node, "roof" = the loaded roof corner of a
    space frame (x = y = 0 at the top), else the tag itself"""
    import numpy as np
    if node is None:
        return int(spec.node_tags[-1])
    if node == "roof":
        c = spec.crd
        top = np.where((c[:, 0] == 0.0) & (c[:, 1] == 0.0) & (c[:, 2] == c[:, 2].max()))[0]
        return int(spec.node_tags[top[0]])
    return int(node)
