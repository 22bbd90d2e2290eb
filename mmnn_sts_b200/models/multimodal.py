"""MultiModalModel with the reference's constructor / attributes / state_dict layout
(/root/reference/models/multimodal.py:9-90).  forward(x) takes the same dict {'image': [B,C,X,Y,Z], 'clinical': [B,F]}
and returns [B,num_classes] or, with `self.blend`, the stacked [3,B,num_classes] (multimodal, image, clinical)."""
import torch
import torch.nn as nn

from ..ops import MLPHeads
from ..utils.utils import BackpropagatableFeatureExtractor
from .mlp import MLP


class MultiModalModel(nn.Module):
    def __init__(self, image_model, clinical_predictors, num_classes, num_features, blend=False):
        super().__init__()
        self.clinical_predictors = clinical_predictors
        self.num_classes = num_classes
        self.num_features = num_features
        self.num_clinical_inputs = len(clinical_predictors)
        # registration order = the reference's (image_model, clinical_model, output_head, clinical_output_head,
        # image_output_head) so that state_dict() enumerates the 779 keys in the same order
        self.image_model = BackpropagatableFeatureExtractor(image_model)
        self.clinical_model = BackpropagatableFeatureExtractor(MLP(self.num_clinical_inputs, self.num_classes, self.num_features))
        self.output_head = nn.Linear(self.num_features * 2, self.num_classes)
        self.blend = blend
        self.clinical_output_head = nn.Linear(self.num_features, self.num_classes)
        self.image_output_head = nn.Linear(self.num_features, self.num_classes)

    def forward(self, x):
        image_data, clinical_data = x["image"], x["clinical"]
        image_features = self.image_model(image_data)
        mlp = self.clinical_model.model
        params, buffers = mlp.kernel_params()
        heads = [self.output_head.weight, self.output_head.bias, self.image_output_head.weight,
                 self.image_output_head.bias, self.clinical_output_head.weight, self.clinical_output_head.bias]
        mask = mlp.sample_masks(clinical_data.shape[0], clinical_data.device)
        preds = MLPHeads.apply(clinical_data, image_features, mask, (self.training, bool(self.blend), self.num_classes),
                               buffers, *params, *heads)
        return preds if self.blend else preds[0]

    @property
    def gradcam_layer(self):
        return self.image_model.model.backbone

    def add_gradcam(self, output_dir):
        """/root/reference/models/multimodal.py:86-90."""
        from ..utils.utils import MultiModalGradCAM
        return MultiModalGradCAM(self)
