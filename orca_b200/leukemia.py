"""
Host-side mirror of the multi-map network classes of /root/reference/orca_leukemia.py, backed by
the same liborca_b200.so modules.

orca_leukemia.py re-declares the orca_modules trees with a `num_2d` parameter -- the number of
Hi-C datasets predicted at once (2 for OrcaLeukemiaA, 6 for OrcaLeukemiaB, :1604-1876):

  Net(num_2d=1, num_1d=None)  :16-509     same as orca_modules.Net, `final` 64 -> max(num_2d,5) -> num_2d
  Decoder(num_2d)             :512-993    combiner inputs 64+num_2d / 128+num_2d channels, nearest x2 upsample
  Decoder_1m(num_2d)          :996-1315
  Encoder()                   :1318-1496  identical to orca_modules.Encoder
  Encoder2()                  :1499-1601  pooling half only, returns [x, d1..d5] (= orca_modules.Encoder2b)

The classes below keep THOSE constructor signatures and state_dict keys (checked against a key
fixture dumped from the reference, tests/golden/state_dict_keys.json) so that
`orca_leukemiaA.*.statedict` files load with strict=True; distenc / coarse / output tensors carry
num_2d channels (`normmats[level]` is (num_2d, 250, 250), orca_predict.py:350).
"""
from . import modules


class Net(modules.Net):
    def __init__(self, num_2d=1, num_1d=None):
        super().__init__(num_1d=num_1d, num_2d=num_2d)


class Decoder(modules.Decoder):
    def __init__(self, num_2d):
        super().__init__(upsample_mode="nearest", num_2d=num_2d)  # nn.Upsample(scale_factor=(2, 2)), :930


class Decoder_1m(modules.Decoder_1m):
    def __init__(self, num_2d):
        super().__init__(num_2d=num_2d)


class Encoder(modules.Encoder):
    pass


class Encoder2(modules.Encoder2b):
    pass
