from .tuber_ava import DETR, PostProcess, PostProcessAVA, build_model  # noqa: F401
