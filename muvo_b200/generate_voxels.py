"""Batch voxeliser: drop-in for ``python data/generate_voxels.py`` (data/generate_voxels.py:110-164).

    python -m muvo_b200.generate_voxels --root <dataset root> [--config data/data_preprocess.yaml] [--io-threads 8]

Walks ``<root>/**/Town*/*/`` exactly as the reference's ``main`` does, reads every run's ``pd_dataframe.pkl``, turns
each (depth_semantic PNG, points_semantic NPY) pair into ``voxel/voxel_<9 digits>.npy`` -- the ``(n,4) uint16``
``[x, y, z, label]`` array of ``voxelize_one`` (data/generate_voxels.py:64-73) -- and stores the relative paths in the
dataframe column ``voxel_path``.  The reference forks ``n_process`` CPU workers; here host threads only decode the
files (cv2 / numpy release the GIL) while the GPU does ``merge_pcd`` + ``voxel_filter`` for ``--batch`` frames per pass
(merged clouds packed back to back on the device, one launch of the point kernels per batch, two host synchronisations
per batch) without a host round trip of the merged clouds.  No hydra / clearml dependency: the YAML is read with ``yaml.safe_load``.
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from types import SimpleNamespace

import numpy as np

DEFAULTS = dict(camera_position=[1.0, 0.0, 2.0], lidar_position=[1.0, 0.0, 2.0], fov=110, voxel_resolution=0.5,
                voxel_size=[192, 192, 64], bev_offset_forward=0, bev_resolution=0.2, offset_z=-20, n_process=8)   # data_preprocess.yaml


def load_config(path=None, **overrides) -> SimpleNamespace:
    cfg = dict(DEFAULTS)
    if path:
        import yaml
        with open(path) as fh:
            cfg.update({k: v for k, v in (yaml.safe_load(fh) or {}).items() if v is not None})
    cfg.update({k: v for k, v in overrides.items() if v is not None})
    return SimpleNamespace(**cfg)


def _read_pair(depth_file: str, lidar_file: str):
    import cv2
    img = cv2.imread(depth_file, -1)
    if img is None:
        raise FileNotFoundError(depth_file)
    data = np.load(lidar_file, allow_pickle=True).item()                          # load_lidar, data_preprocessing.py:80-84
    return img, data['points_xyz'], data['ObjTag']


def voxelize_frame(img, lidar_xyz, lidar_sem, cfg) -> np.ndarray:
    """merge_pcd + voxel_filter on the GPU for one decoded frame -> ``(n,4) uint16`` (generate_voxels.py:64-70)."""
    from .points import GridSpec, merge_pcd_device, sensor_to_grid
    xyz, sem = merge_pcd_device(img, lidar_xyz, lidar_sem, cfg.camera_position, cfg.lidar_position, cfg.fov)
    offset_x = cfg.bev_offset_forward * cfg.bev_resolution
    offset_z = cfg.offset_z * cfg.voxel_resolution
    spec = GridSpec(cfg.voxel_resolution, tuple(cfg.voxel_size), (offset_x, 0, offset_z))
    r = sensor_to_grid(xyz, sem, None, grid=spec, dense=False, sparse=True)
    n = int(r["n_occ"][0].item())
    return r["voxel_sparse"][:n].cpu().numpy().view(np.uint16)


def voxelize_frames(frames, cfg) -> list:
    """``voxelize_frame`` for a list of decoded frames in ONE pass: N stream-ordered merges packed back to back on the device,
    one launch of the point kernels over the ragged batch, two host synchronisations in total (the row offsets, then the
    packed sparse lists) instead of two per frame.  Returns the N ``(n,4) uint16`` arrays."""
    from .points import GridSpec, merge_pcd_batch, sensor_to_grid
    if not frames:
        return []
    xyz, sem, offs = merge_pcd_batch(frames, cfg.camera_position, cfg.lidar_position, cfg.fov)
    total = int(offs[-1].item())                                                   # sync 1: size of the merged batch
    offset_x = cfg.bev_offset_forward * cfg.bev_resolution
    offset_z = cfg.offset_z * cfg.voxel_resolution
    spec = GridSpec(cfg.voxel_resolution, tuple(cfg.voxel_size), (offset_x, 0, offset_z))
    r = sensor_to_grid(xyz[:total], sem[:total], offs, grid=spec, dense=False, sparse=True, packed_sparse=True)
    start = r["sparse_start"].cpu().numpy()                                        # sync 2: rows of every frame
    rows = r["voxel_sparse"][:int(start[-1])].cpu().numpy().view(np.uint16)
    return [np.ascontiguousarray(rows[start[f]:start[f + 1]]) for f in range(len(frames))]


def voxelize_run(run_dir: Path, cfg, io_threads: int = 8, progress=None, batch: int = 16) -> int:
    """One run directory (``.../TownXX/NNNN/``): generate_voxels.py:126-163.  Returns the number of frames written."""
    import pandas as pd
    pd_file = run_dir / 'pd_dataframe.pkl'
    frame = pd.read_pickle(pd_file)
    save_path = run_dir / 'voxel'
    if save_path.exists():
        shutil.rmtree(save_path)
    save_path.mkdir()
    jobs = []
    for j in range(len(frame)):
        row = frame.iloc[j]
        depth_file = str(run_dir / row['depth_semantic_path'])
        lidar_file = str(run_dir / row['points_semantic_path'])
        name = re.match(r'.*/.*_(\d{9})\.png', depth_file).group(1)
        name_ = re.match(r'.*/.*_(\d{9})\.npy', lidar_file).group(1)
        assert name == name_, 'file sequence is false.'
        jobs.append((depth_file, lidar_file, f'{save_path.name}/voxel_{name}.npy'))
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        ahead = max(2 * io_threads, 2 * batch)
        pending = [pool.submit(_read_pair, d, l) for d, l, _ in jobs[:ahead]]
        nxt = len(pending)
        savers = []
        for j0 in range(0, len(jobs), max(1, batch)):                              # `batch` frames per merge + voxelise pass
            j1 = min(len(jobs), j0 + max(1, batch))
            decoded = []
            for j in range(j0, j1):
                decoded.append(pending[j].result())
                pending[j] = None
                if nxt < len(jobs):                                                # keep the decoders busy
                    pending.append(pool.submit(_read_pair, jobs[nxt][0], jobs[nxt][1]))
                    nxt += 1
            for (_, _, rel), arr in zip(jobs[j0:j1], voxelize_frames(decoded, cfg)):
                savers.append(pool.submit(np.save, str(run_dir / rel), arr))        # file writes overlap the next batch
            if progress is not None:
                progress(j1 - j0)
        for sv in savers:
            sv.result()
    frame['voxel_path'] = [rel for _, _, rel in jobs]
    frame.to_pickle(pd_file)
    return len(jobs)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--root", required=True)
    ap.add_argument("--config", default=None, help="data_preprocess.yaml (defaults = the reference's values)")
    ap.add_argument("--io-threads", type=int, default=None)
    ap.add_argument("--batch", type=int, default=16, help="frames per merge + voxelise pass on the GPU")
    a = ap.parse_args(argv)
    cfg = load_config(a.config)
    root = Path(a.root)
    runs = sorted(p for p in root.glob('**/Town*/*/') if p.is_dir())              # generate_voxels.py:118
    if not root.exists() or not runs:
        print('Root Path does not EXIST or there are NO LEGAL files!!!')
        return 1
    io = a.io_threads if a.io_threads is not None else int(cfg.n_process)
    for i, run in enumerate(runs):
        n = voxelize_run(run, cfg, io, batch=a.batch)
        print(f'{i + 1}/{len(runs)} {run}: {n} frames -> {run / "voxel"}')
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
