"""Tile binning entry point (mirror of splat_py/tile_culling.py:8-27)."""
from __future__ import annotations

from . import native

MAX_TILES_PER_GAUSSIAN = 1024  # accepted and ignored by the native op, exactly as in the reference


def get_splats(uvs, tiles, conic, xyz_camera_frame, mh_dist):
    """(gaussian, tile) pairs of the mh_dist-sigma footprints, ordered by (tile, camera depth, gaussian index).

    Returns ``(sorted_gaussian_idx_by_splat_idx int32 [P], splat_start_end_idx_by_tile_idx int32 [tiles + 1])``:
    the splats of tile t are ``sorted[ranges[t]:ranges[t + 1]]``.  One host sync (to size the outputs).
    Unlike the reference there is no process exit on non-finite positions: such points sort last in their tile.
    """
    nx, ny = tiles.x_tiles_count, tiles.y_tiles_count
    return native().get_sorted_gaussian_list(MAX_TILES_PER_GAUSSIAN, uvs, xyz_camera_frame, conic, nx, ny, mh_dist)
