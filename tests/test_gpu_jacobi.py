"""General-covariance MUSIC (rs_music_covariance): the Hermitian Jacobi eigensolver against numpy.linalg.eigh and the
noise-subspace pseudo-spectrum against a numpy restatement of angle_estimation.py:127-152 for multi-snapshot
covariances and num_sources >= 1; and its agreement with the closed-form path on the reference's rank-1 case."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu


def _pipe(A, res=1.0):
    from radar_slam_b200 import RadarConfig, FramePipeline
    return FramePipeline(RadarConfig(num_antennas=A, search_resolution=res))


def _music_numpy(R, steer, K):
    vals, vecs = np.linalg.eigh(R)
    En = vecs[:, ::-1][:, K:]
    T = steer.conj() @ En
    den = np.abs(np.sum(T * T.conj(), axis=1))
    with np.errstate(divide="ignore"):
        return vals[::-1], np.where(den > 1e-12, 1.0 / den, 0.0)


@pytest.mark.parametrize("A,K,snaps", [(8, 1, 24), (8, 2, 24), (4, 1, 6), (16, 3, 40), (5, 2, 12), (32, 4, 64), (2, 1, 4)])
def test_jacobi_eigen_and_music_spectrum(A, K, snaps):
    rng = np.random.RandomState(A * 100 + K)
    p = orc.RadarParams(num_antennas=A)
    grid = orc.azimuth_grid((-90, 90), 1.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    n = 64
    covs = np.zeros((n, A, A), dtype=np.complex128)
    truth = []
    for i in range(n):
        # K sources at random on-grid-ish angles + noise, `snaps` snapshots
        ang = rng.choice(np.arange(20, len(grid) - 20), size=K, replace=False)
        amp = (rng.randn(K, snaps) + 1j * rng.randn(K, snaps)) * rng.uniform(1.0, 3.0, (K, 1))
        X = steer[ang].T @ amp + 0.3 * (rng.randn(A, snaps) + 1j * rng.randn(A, snaps))
        covs[i] = X @ X.conj().T / snaps
        truth.append(ang)
    pipe = _pipe(A)
    out = pipe.music_covariance(torch.from_numpy(covs.astype(np.complex64)).cuda(), num_sources=K, want_vectors=True)
    torch.cuda.synchronize()
    vals = out["eigvals"].cpu().numpy()
    vecs = out["eigvecs"].cpu().numpy().astype(np.complex128)
    spec = out["spectrum"].cpu().numpy()
    for i in range(n):
        R = covs[i].astype(np.complex64).astype(np.complex128)
        w, sp = _music_numpy(R, steer, K)
        scale = w[0]
        assert np.abs(vals[i] - w).max() < 2e-5 * scale                              # eigenvalues, descending
        V = vecs[i]
        assert np.abs(V.conj().T @ V - np.eye(A)).max() < 2e-5                       # unitary
        assert np.abs(R @ V - V * vals[i][None, :]).max() < 5e-5 * scale             # eigen-equation
        # pseudo-spectrum: compare denominators (the spectrum itself is 1/den and huge at the peaks)
        den_gpu, den_ref = 1.0 / np.maximum(spec[i], 1e-30), 1.0 / np.maximum(sp, 1e-30)
        assert np.abs(den_gpu - den_ref).max() < 2e-4 * A
        # same peak as the numpy pseudo-spectrum unless its top two denominators are within fp32 resolution
        srt = np.sort(den_ref)
        if srt[1] - srt[0] > 1e-3 * A:
            assert int(out["aidx"][i]) == int(np.argmax(sp))


def test_rank_one_case_matches_closed_form_path():
    """For the reference's single-snapshot covariance R = s s^H the general path and the closed form used by the
    batched kernels give the same argmax (outside fp32 ties)."""
    A = 8
    rng = np.random.RandomState(5)
    p = orc.RadarParams(num_antennas=A)
    grid = orc.azimuth_grid((-90, 90), 1.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    s = rng.randn(200, A) + 1j * rng.randn(200, A)
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    covs = s[:, :, None] * s.conj()[:, None, :]
    pipe = _pipe(A)
    out = pipe.music_covariance(torch.from_numpy(covs.astype(np.complex64)).cuda(), num_sources=1)
    spec = orc.beamforming_spectra(s, steer)
    idx = np.argmax(spec, axis=1)
    srt = np.sort(spec, axis=1)
    gap = (srt[:, -1] - srt[:, -2]) / srt[:, -1]
    bad = out["aidx"].cpu().numpy() != idx
    assert np.all(gap[bad] < 1e-4) and bad.mean() < 0.05
    vals = out["eigvals"].cpu().numpy()
    assert np.abs(vals[:, 0] - 1.0).max() < 1e-5 and np.abs(vals[:, 1:]).max() < 1e-5
