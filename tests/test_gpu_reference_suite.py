"""GPU: the reference's OWN unit tests (test/test_*.c, compiled unmodified in the build container by
`make -C oracle dropin-tests` against oracle/shim/check.h) linked against libsdrmodem_b200.so instead of the reference's
src/dsp. This is the drop-in proof: same sources, same assertions, same fixtures, GPU library underneath.
The binaries live in oracle/_ref/tests (git-ignored, they travel to the GPU box) and open fixtures by bare name."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

BIN_DIR = os.path.join(ROOT, "oracle", "_ref", "tests")
TESTS = ["test_lpf", "test_lpf_taps", "test_quadrature_demod", "test_dc_blocker", "test_clock_recovery_mm",
         "test_mmse_fir_interpolator", "test_sig_source", "test_gaussian_taps", "test_interp_fir_filter",
         "test_frequency_modulator", "test_gfsk_mod", "test_fsk_demod", "test_doppler", "test_queue",
         # a caller of the path: the reference's file SDR plugin (src/sdr/file_source.c) on the product's sig_source
         "test_file_source",
         # dsp_worker_create with the reference's signature; requests built by the reference's test/utils.c through the
         # product's protobuf codec (make -C oracle server-tests)
         "test_dsp_worker",
         # the reference's whole control plane, unmodified (src/tcp_server.c, sdr_worker.c, server_config.c, SDR plugins, its
         # test client and SDR mocks), on top of the product library: RX and TX sessions over real sockets
         "test_tcp_server"]


@pytest.mark.parametrize("name", TESTS)
def test_reference_test_binary_passes_against_the_gpu_library(name):
    binary = os.path.join(BIN_DIR, name)
    if not os.path.exists(binary):
        pytest.skip("%s was not built (needs /root/reference at build time)" % binary)
    # SDRM_WARM_START: the CUDA context is created when the library is loaded, not inside the first client's request (the
    # reference's server tests give a client two seconds for its response; context creation alone takes about that long)
    proc = subprocess.run([binary], cwd=GOLDEN, capture_output=True, text=True, timeout=300, env=dict(os.environ, SDRM_WARM_START="0"))
    tail = (proc.stdout + proc.stderr)[-2000:]
    assert proc.returncode == 0, tail
    assert "0 failed" in proc.stdout, tail


def test_reference_perf_program_runs_against_the_gpu_library():
    """test/perf_fsk_modem.c, the program behind every number the reference publishes, unmodified, on the single-handle
    drop-in path (100 x 4096-sample fsk_demod_process calls and 100 x 2048-byte gfsk_mod_process calls per figure)."""
    binary = os.path.join(BIN_DIR, "perf_fsk_modem")
    if not os.path.exists(binary):
        pytest.skip("%s was not built (needs /root/reference at build time)" % binary)
    proc = subprocess.run([binary], cwd=GOLDEN, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, (proc.stdout + proc.stderr)[-2000:]
    lines = [line for line in proc.stdout.splitlines() if "completed" in line]
    assert len(lines) == 2, proc.stdout
    for line in lines:
        seconds = float(line.split(":")[1].split()[0])
        assert 0.0 < seconds < 5.0, line
    print("\n".join(lines))
