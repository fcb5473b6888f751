"""Layer tables of the three graphs on the hot path, shared by the zoo, the Python-side program
assembly and the file formats.  (csrc/xemo_net.cu holds the same tables for the graph-level C ABI: the library does not
depend on this package; tests/test_host.py::test_layer_tables_agree_with_the_library checks them against each other.)"""

# VGGVox student: name, FH, FW, Cin, Cout, stride, pad, has_bn
STUDENT_CONVS = [
    ("conv1", 7, 7, 1, 96, (2, 2), (1, 1, 1, 1), True),
    ("conv2", 5, 5, 96, 256, (2, 2), (1, 1, 1, 1), True),
    ("conv3", 3, 3, 256, 384, (1, 1), (1, 1, 1, 1), True),
    ("conv4", 3, 3, 384, 256, (1, 1), (1, 1, 1, 1), True),
    ("conv5", 3, 3, 256, 256, (1, 1), (1, 1, 1, 1), True),
    ("fc6", 9, 1, 256, 4096, (1, 1), (0, 0, 0, 0), True),
    ("fc7", 1, 1, 4096, 1024, (1, 1), (0, 0, 0, 0), True),
    ("fc8", 1, 1, 1024, 8, (1, 1), (0, 0, 0, 0), False),
]
STUDENT_POOLS = {"conv1": ("max", (3, 3), (2, 2)), "conv2": ("max", (3, 3), (2, 2)), "conv5": ("max", (5, 3), (3, 2)),
                 "fc6": ("avg", None, (1, 1))}
TEACHER_STAGES = [(3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)]
