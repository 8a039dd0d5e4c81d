"""Tile binning entry point (mirror of splat_py/tile_culling.py:8-27)."""
from __future__ import annotations

from . import native


def get_splats(uvs, tiles, conic, xyz_camera_frame, mh_dist):
    """-> (sorted_gaussian_idx_by_splat_idx int32 [P], splat_start_end_idx_by_tile_idx int32 [T+1]).

    The reference aborts the process on non-finite xyz (a host sync); here non-finite depths
    simply sort last inside their tile.
    """
    return native().get_sorted_gaussian_list(
        1024,  # max_tiles_per_gaussian: accepted and ignored, as in the reference
        uvs,
        xyz_camera_frame,
        conic,
        tiles.x_tiles_count,
        tiles.y_tiles_count,
        mh_dist,
    )
