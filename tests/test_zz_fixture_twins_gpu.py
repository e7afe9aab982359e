"""CUDA-path twins of the reference-fixture checks in tests/fixture_checks.py (BASELINE.json configs[0] and [1] at their
real sizes, get_likelihood, the transformer sampler).  The same bodies run on the CPU stand-in in
tests/test_modules_cpu.py.  These twins were written after round 1's GPU budget was spent — the file sorts last so that
`pytest -x` reaches every previously verified GPU test first."""
import pytest

pytestmark = pytest.mark.gpu


def test_rank1_rank3_reference_fixtures(cuda_device, monkeypatch):
    """get_likelihood (12 DDPM steps, two prediction types, KL maps) and the transformer forward / greedy
    VQVAETransformerInferer.sample against fixtures written by the unmodified reference (make_golden_next.py); the same
    bodies run on the CPU stand-in in tests/test_modules_cpu.py."""
    from tests import fixture_checks
    fixture_checks.check_likelihood_fixture("cuda", monkeypatch)
    fixture_checks.check_transformer_fixture("cuda")


def test_c1_reference_fixture(cuda_device):
    """BASELINE.json configs[0]: tutorial 2-D UNet (128, 256, 256), DDPM with 4 inference steps, batch 2 of 1x64x64 —
    the CUDA path against the unmodified reference's CPU run (tests/golden/g_c1.pt)."""
    from tests import fixture_checks
    fixture_checks.check_c1_fixture("cuda")


def test_c2_reference_fixture(cuda_device, capsys):
    """BASELINE.json configs[1]: LDM-tutorial AutoencoderKL + latent UNet (DDIM-50 trajectory of the unmodified
    reference, teacher-forced at four probe steps) and the decoder, on the CUDA path.  The ill-conditioned probe
    (t = 500, see tests/fixture_checks.py) is reported, and bounded only on the CPU stand-in where it was measured."""
    from tests import fixture_checks
    report = fixture_checks.check_c2_fixture("cuda", strict_ill_conditioned=False)
    with capsys.disabled():
        print(f"\n[C2 probes] relative L2 (network output, next latent): {report}")
