"""TEST INFRASTRUCTURE ONLY -- empty stand-in for pyfaidx."""


class Fasta:
    def __init__(self, *a, **k):
        raise NotImplementedError
