class Matcher(object):
    """training-only; constructed by the EMM target sampler factory, never called at inference."""

    def __init__(self, *a, **k):
        pass
