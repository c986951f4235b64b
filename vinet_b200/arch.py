"""Architecture tables of the ViNet / AViNet hot path (product-owned copy).

Only *numbers* (channel counts, kernel sizes) read off the reference; consumed by the plan builder
and the parameter-holder modules (both in ``vinet_b200/model.py`` / ``avmodel.py``).

Reference sites:
  * Inception ("Mixed") blocks ........ model_utils.py:162-420
  * S3D backbone stage layout ......... model.py:690-743
  * decoders by clip length ........... model.py:251-311 (T=32), :313 (T=16), :375 (T=8), :437 (T=48)
  * SoundNet .......................... model.py:746-825
"""

# name -> (Cin, b0, b1_reduce, b1_out, b2_reduce, b2_out, b3_out)       model_utils.py:162-420
MIXED = {
    "3b": (192, 64, 96, 128, 16, 32, 32),
    "3c": (256, 128, 128, 192, 32, 96, 64),
    "4b": (480, 192, 96, 208, 16, 48, 64),
    "4c": (512, 160, 112, 224, 24, 64, 64),
    "4d": (512, 128, 128, 256, 24, 64, 64),
    "4e": (512, 112, 144, 288, 32, 64, 64),
    "4f": (528, 256, 160, 320, 32, 128, 128),
    "5b": (832, 256, 160, 320, 32, 128, 128),
    "5c": (832, 384, 192, 384, 48, 128, 128),
}

# backbone stages: attribute name -> list of Mixed names                  model.py:699-718
STAGES = {"base2": ["3b", "3c"], "base3": ["4b", "4c", "4d", "4e", "4f"], "base4": ["5b", "5c"]}


def mixed_out(name):
    c = MIXED[name]
    return c[1] + c[3] + c[5] + c[6]


# Decoder tail (the part of ``convtsp4`` after the first conv+ReLU+upsample) by clip length.
# Each entry: list of Sequential items  ('conv', Cin, Cout, (kt,kh,kw), (st,1,1), pad_hw, bias) |
# 'relu' | 'up' | 'sigmoid'.  Indices in the nn.Sequential follow the list position + 3.
#                                                                         model.py:272-283, 334-345, 396-407, 458-469
def decoder_tail(num_clips):
    if num_clips == 32:
        return [("conv", 64, 32, (2, 3, 3), (2, 1, 1), 1, False), "relu", "up",
                ("conv", 32, 32, (2, 1, 1), (2, 1, 1), 0, False), "relu",
                ("conv", 32, 1, (1, 1, 1), (1, 1, 1), 0, True), "sigmoid"]
    if num_clips == 48:
        return [("conv", 64, 32, (2, 3, 3), (2, 1, 1), 1, False), "relu", "up",
                ("conv", 32, 32, (3, 1, 1), (3, 1, 1), 0, True), "relu",
                ("conv", 32, 1, (1, 1, 1), (1, 1, 1), 0, True), "sigmoid"]
    if num_clips == 16:
        return [("conv", 64, 32, (2, 3, 3), (2, 1, 1), 1, False), "relu", "up",
                ("conv", 32, 1, (1, 1, 1), (1, 1, 1), 0, True), "sigmoid"]
    if num_clips == 8:
        return [("conv", 64, 32, (1, 3, 3), (1, 1, 1), 1, False), "relu", "up",
                ("conv", 32, 1, (1, 1, 1), (1, 1, 1), 0, True), "sigmoid"]
    raise ValueError("the reference has a decoder only for num_clips in {8,16,32,48} (model.py:92-99)")


# Decoder head convs (shared by all clip lengths): (Cin, Cout, kt)        model.py:256-271
DECODER_HEAD = [(1024, 832, 1), (832, 480, 3), (480, 192, 5), (192, 64, 5)]


def decoder_head(num_hier=3):
    """Head convs of the decoder that fuses `num_hier` skip levels (DecoderConvUpNoHier / 1Hier / 2Hier / DecoderConvUp,
    model.py:501,564,627,251): stage i > num_hier has no skip tensor to collapse, so its kernel is (1,3,3)."""
    return [(cin, cout, kt if i <= num_hier else 1) for i, (cin, cout, kt) in enumerate(DECODER_HEAD)]

# SoundNet layers 1..7: (Cin, Cout, k, pad, pool)  stride is always 2     model.py:750-786
SOUNDNET = [(1, 16, 64, 32, 8), (16, 32, 32, 16, 8), (32, 64, 16, 8, 1), (64, 128, 8, 4, 1),
            (128, 256, 4, 2, 4), (256, 512, 4, 2, 1), (512, 1024, 4, 2, 1)]
AUDIO_LEN = 70560
