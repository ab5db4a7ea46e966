"""B200-native top-down pose hot path (frame preprocess -> detector-head decode +
NMS -> person crop -> HRNet -> heatmap decode + body-proportion lengths).

Drop-in Python surface of SamSamhuns/human_body_proportion_estimation for that
path; every stage is a hand-written sm_100a kernel behind the C ABI in
include/hbp.h (libhbp_b200.so).  Importing the package never touches the GPU;
creating an Engine does, and fails loudly without one.
"""
__all__ = ["Engine", "MultiGpuEngine", "PoseEstimator", "run_pdet_pose", "run_demo_pose_est",
           "detect_onnx", "non_max_suppression", "w_non_max_suppression", "letterbox_image",
           "scale_coords"]


def __getattr__(name):
    if name in ("Engine", "MultiGpuEngine"):
        from . import engine
        return getattr(engine, name)
    if name == "PoseEstimator":
        from .pose_estimator import PoseEstimator
        return PoseEstimator
    if name == "run_pdet_pose":
        from .person_det_pose import run_pdet_pose
        return run_pdet_pose
    if name == "run_demo_pose_est":
        from .pose_est_hrnet import run_demo_pose_est
        return run_demo_pose_est
    if name == "detect_onnx":
        from .obj_det_yolov5 import detect_onnx
        return detect_onnx
    if name in ("non_max_suppression", "w_non_max_suppression", "letterbox_image", "scale_coords"):
        from . import onnx_utils
        return getattr(onnx_utils, name)
    raise AttributeError(name)
