"""bbpcg.synth.cages_torch (what bench.py --parts builds the 1000-sphere case with, on the GPU) against the CPU oracle's
restatement of cuda_build_cages (oracle/pcg_ref.c: bbo_build_cages; src/cuda_particle.cu:1516-1646,
src/particle_kernel.cu:135-576): phase, phase_shell and the three flag arrays must agree bit for bit on every block."""
import numpy as np
import pytest

from bbpcg import synth
from cases import Case
from oracle import binding as ob


@pytest.mark.parametrize("cells,blocks,bc,nparts,radius", [((40, 40, 40), (1, 1, 1), "sedimentation", 4, 2.5),
                                                          ((33, 28, 44), (2, 1, 2), "duct", 5, 1.3),
                                                          ((24, 24, 24), (1, 2, 1), "box", 3, 1.9),
                                                          ((32, 32, 32), (1, 1, 1), "periodic", 6, 1.5)])
def test_cages_torch_matches_the_oracle(cells, blocks, bc, nparts, radius):
    case = Case(cells, blocks=blocks, bc=bc, nparts=nparts, radius=radius)
    solid = 0
    for r in range(case.o.nblocks):
        ph, sh, fu, fv, fw = synth.cages_torch(case.o.dom(r), case.o.DOM, case.o.bc, case.parts, "cpu")
        a = case.o.array
        assert np.array_equal(ph.numpy(), a(r, ob.PHASE))
        assert np.array_equal(sh.numpy(), a(r, ob.PHASE_SHELL))
        assert np.array_equal(fu.numpy(), a(r, ob.FLAG_U))
        assert np.array_equal(fv.numpy(), a(r, ob.FLAG_V))
        assert np.array_equal(fw.numpy(), a(r, ob.FLAG_W))
        solid += int((a(r, ob.PHASE)[1:-1, 1:-1, 1:-1] > -1).sum())
    assert solid > 50 * nparts / 2


def test_random_spheres_config_c4_density():
    """BASELINE configs[3]: 1000 non-overlapping spheres of 8 cells per radius fit a 512^3 box of L = 64"""
    from bbpcg.grid import DomStruct
    DOM = DomStruct()
    DOM.xs = DOM.ys = DOM.zs = 0.
    DOM.xe = DOM.ye = DOM.ze = 64.
    DOM.dx = DOM.dy = DOM.dz = 64. / 512
    x, y, z, r = synth.random_spheres(DOM, 1000, 1.0)
    pts = np.stack([x, y, z], 1)
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1) + np.eye(1000) * 1e9
    assert d2.min() > (2.2 * 1.0) ** 2 and len(x) == 1000
