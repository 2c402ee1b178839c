from sparse_coding_b200.topk_encoder import TopKEncoder, TopKLearnedDict  # noqa: F401
