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
    reference, teacher-forced at four probe steps — including the ill-conditioned t = 500 one) and the decoder, on the
    CUDA path, all held to the suite's tolerance (tests/fixture_checks.py: TOL_REL / TOL_MAX)."""
    from tests import fixture_checks
    report = fixture_checks.check_c2_fixture("cuda")
    with capsys.disabled():
        print(f"\n[C2] (relative L2, normalised max-abs): {report}")


def test_c3_reference_fixture(cuda_device, capsys):
    """BASELINE.json configs[2], the bench's model (3-D UNet (256, 256, 512), head 512) at 32x40x32: forward + DDIM-5
    against the unmodified reference (tests/golden/g_c3.pt)."""
    from tests import fixture_checks
    report = fixture_checks.check_c3_fixture("cuda")
    with capsys.disabled():
        print(f"\n[C3] (relative L2, normalised max-abs): {report}")


def test_c4_reference_fixture(cuda_device, capsys):
    """BASELINE.json configs[3]: VQVAE (256, 256) / 256 codes x 32 at 64^3 — indices bit-exact on the reference's
    encoder output, flips through this encoder only at the reference's near-ties (tests/golden/g_c4.pt)."""
    from tests import fixture_checks
    report = fixture_checks.check_c4_fixture("cuda")
    with capsys.disabled():
        print(f"\n[C4] {report}")


def test_c5_reference_fixture(cuda_device, capsys):
    """BASELINE.json configs[4]: ControlNet + conditioned UNet, one classifier-free-guidance DDIM step at 3x256x256
    (T = 16 384) against the unmodified reference (tests/golden/g_c5.pt)."""
    from tests import fixture_checks
    report = fixture_checks.check_c5_fixture("cuda")
    with capsys.disabled():
        print(f"\n[C5] (relative L2, normalised max-abs): {report}")
