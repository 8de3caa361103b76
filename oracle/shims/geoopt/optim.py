"""train.py:283-288 of the reference names geoopt.optim.RiemannianAdam; training is out of scope,
the name only has to exist so that `import train` does not fail."""
import torch


class RiemannianAdam(torch.optim.Adam):
    pass
