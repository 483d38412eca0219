"""`import pointnet2_cuda` drop-in: put this directory on sys.path in place of the reference's
compiled extension (model/pointnet2_utils.py:7 imports it by this name)."""
from pn2_b200.pointnet2_cuda import *  # noqa: F401,F403
from pn2_b200.pointnet2_cuda import (ball_query_wrapper, furthest_point_sampling_wrapper,  # noqa: F401
                                     gather_points_grad_wrapper, gather_points_wrapper, group_points_grad_wrapper,
                                     group_points_wrapper, three_interpolate_grad_wrapper, three_interpolate_wrapper,
                                     three_nn_wrapper)
