import torch

from maskrcnn_benchmark.structures.boxlist_ops import cat_boxlist


class RPNPostProcessor(torch.nn.Module):
    """Upstream per-level select + cross-level top-k.  forward_for_single_feature_map is
    overridden by siammot/operator_patch/rpn_patch.py:15-60."""

    def __init__(self, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None,
                 fpn_post_nms_top_n=None, fpn_post_nms_per_batch=True):
        super().__init__()
        self.pre_nms_top_n = pre_nms_top_n
        self.post_nms_top_n = post_nms_top_n
        self.nms_thresh = nms_thresh
        self.min_size = min_size
        self.box_coder = box_coder
        self.fpn_post_nms_top_n = post_nms_top_n if fpn_post_nms_top_n is None else fpn_post_nms_top_n
        self.fpn_post_nms_per_batch = fpn_post_nms_per_batch

    def forward_for_single_feature_map(self, anchors, objectness, box_regression):
        raise NotImplementedError("patched by siammot.operator_patch.rpn_patch")

    def forward(self, anchors, objectness, box_regression, targets=None):
        sampled = []
        num_levels = len(objectness)
        anchors = list(zip(*anchors))
        for a, o, b in zip(anchors, objectness, box_regression):
            sampled.append(self.forward_for_single_feature_map(a, o, b))
        boxlists = [cat_boxlist(list(bl)) for bl in zip(*sampled)]
        if num_levels > 1:
            boxlists = self.select_over_all_levels(boxlists)
        return boxlists

    def select_over_all_levels(self, boxlists):
        assert not self.training
        for i in range(len(boxlists)):
            objectness = boxlists[i].get_field("objectness")
            k = min(self.fpn_post_nms_top_n, len(objectness))
            _, inds = torch.topk(objectness, k, dim=0, sorted=True)
            boxlists[i] = boxlists[i][inds]
        return boxlists


def make_rpn_postprocessor(config, rpn_box_coder, is_train):
    raise NotImplementedError("patched by siammot.operator_patch.rpn_patch")
