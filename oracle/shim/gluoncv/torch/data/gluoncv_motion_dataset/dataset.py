"""Stand-in for the two gluoncv data classes the reference's result egress uses (siammot/utils/boxlists_to_entities.py:3,
siammot/engine/inferencer.py:134-153): plain containers with the members those call sites touch.  TEST INFRASTRUCTURE
(gluoncv is not installable offline); see oracle/shim/maskrcnn_benchmark/__init__.py."""


class AnnoEntity(object):
    def __init__(self, time=-1, id=-1):
        self.time = time
        self.id = id
        self.frame_num = None
        self.bbox = None
        self.confidence = 1.0
        self.labels = None


class DataSample(object):
    """A video sample: metadata (width / height / fps), entities, and -- for the inference loader
    (build_inference_data_loader.py:21-40) -- a frame reader whose items are (PIL image, timestamp, extra)."""

    def __init__(self, id="sample", raw_info=None, metadata=None, entities=None, frames=None):
        self.id = id
        self.raw_info = raw_info
        self.metadata = dict(metadata or {})
        self.entities = []
        self._frames = list(frames or [])
        for e in entities or []:
            self.add_entity(e)

    @property
    def width(self):
        return self.metadata["resolution"]["width"]

    @property
    def height(self):
        return self.metadata["resolution"]["height"]

    def __len__(self):
        return len(self._frames)

    def get_data_reader(self):
        fps = float(self.metadata.get("fps", 30.0))
        return [(im, int(1000.0 * i / fps), None) for i, im in enumerate(self._frames)]

    def get_entities_for_frame_num(self, frame_num):
        return [e for e in self.entities if e.frame_num == frame_num]

    def add_entity(self, entity):
        self.entities.append(entity)

    def get_entities_with_id(self, id):
        return [e for e in self.entities if e.id == id]

    def get_copy_without_entities(self):
        return DataSample(self.id, metadata=self.metadata)
