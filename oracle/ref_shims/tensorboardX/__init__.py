"""Stand-in for tensorboardX (absent from the image): scripts/hdqn.py:12 imports SummaryWriter at module
level; golden-vector generation never logs.  TEST INFRASTRUCTURE ONLY."""


class SummaryWriter:
    def __init__(self, *args, **kwargs):
        pass

    def add_scalar(self, *args, **kwargs):
        pass

    def close(self):
        pass
