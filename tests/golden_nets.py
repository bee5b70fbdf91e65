"""Net specs shared by the golden generator and the tests (oracle syntax)."""
NETS = {
    "mlp_sigmoid_softmax_cxent_sgd": (11, (20,), [("dense", 12, "sigmoid"), ("dense", 5, "softmax")], "cxent", "sgd", 7),
    "mlp_tanh_linear_mse_momentum": (12, (9,), [("dense", 8, "tanh"), ("dense", 6, "relu"), ("dense", 4, "linear")], "mse", "momentum", 5),
    "lenet_sgd": (13, (1, 10, 10), [("conv", 4, 3, 3, "linear"), ("pool", 2), ("act", "relu"), ("flatten",),
                                    ("dense", 5, "softmax")], "cxent", "sgd", 6),
    "conv3_adam": (14, (3, 30, 30), [("conv", 4, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                     ("conv", 6, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                     ("conv", 8, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                     ("flatten",), ("dense", 10, "softmax")], "cxent", "adam", 3),
    "conv_tanh_bias_after_act": (15, (2, 8, 8), [("conv", 3, 3, 3, "tanh"), ("flatten",), ("dense", 4, "softmax")],
                                 "cxent", "sgd", 4),
    "lstm_dense_sgd": (16, (6, 5), [("lstm", 8, "tanh", False), ("dense", 4, "softmax")], "cxent", "sgd", 5),
    "lstm_seq_stack_adam": (17, (4, 3), [("lstm", 6, "tanh", True), ("lstm", 5, "tanh", False), ("dense", 3, "linear")],
                            "mse", "adam", 4),
}
