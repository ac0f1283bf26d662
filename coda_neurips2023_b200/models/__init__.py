"""`models.build_model(args, dataset_config) -> (model, BoxProcessor)` with the
reference's registry keys (models/__init__.py:3-10)."""
from .model_3detr import build_3detr_multiclasshead, build_3detr_predictedbox_distillation_head

MODEL_FUNCS = {
    "3detrmulticlasshead": build_3detr_multiclasshead,
    "3detr_predictedbox_distillation": build_3detr_predictedbox_distillation_head,
}


def build_model(args, dataset_config):
    model, processor = MODEL_FUNCS[args.model_name](args, dataset_config)
    return model, processor
