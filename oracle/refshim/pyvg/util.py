"""TEST INFRASTRUCTURE ONLY."""


def get_stats(*a, **k):
    raise NotImplementedError
