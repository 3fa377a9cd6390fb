"""TEST INFRASTRUCTURE: import-only stand-in."""


class RecurrentPPO:
    @staticmethod
    def load(*a, **k):
        raise NotImplementedError("shim: RecurrentPPO is import-only here")
