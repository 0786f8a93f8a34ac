"""Host-side mirror of the reference's ``utils`` package (depth_operations.py, dense_image_warp.py)."""
from .dense_image_warp import dense_image_warp, back_project, back_project_grad
from .depth_operations import (get_rot_mat, get_coords_2d, parallax2depth, depth2parallax, prev_d2para,
                               get_parallax_sweeping_cv, get_parallax_sweeping_cv_grad, cost_volume, tile_in_batch)
