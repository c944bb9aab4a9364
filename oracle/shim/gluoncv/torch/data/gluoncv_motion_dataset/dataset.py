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
    def __init__(self, id="sample", entities=None, metadata=None):
        self.id = id
        self.metadata = dict(metadata or {})
        self.entities = []
        for e in entities or []:
            self.add_entity(e)

    def add_entity(self, entity):
        self.entities.append(entity)

    def get_entities_with_id(self, id):
        return [e for e in self.entities if e.id == id]

    def get_copy_without_entities(self):
        return DataSample(self.id, metadata=self.metadata)
