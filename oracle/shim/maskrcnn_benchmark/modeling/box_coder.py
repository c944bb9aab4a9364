from oracle import prims


class BoxCoder(object):
    def __init__(self, weights, bbox_xform_clip=prims.BBOX_XFORM_CLIP):
        self.weights = weights
        self.bbox_xform_clip = bbox_xform_clip

    def decode(self, rel_codes, boxes):
        return prims.box_decode(rel_codes, boxes, self.weights)
