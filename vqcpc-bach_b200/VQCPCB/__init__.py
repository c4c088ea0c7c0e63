"""vqcpc_b200: B200-native VQ-CPC encoder-training hot path behind the reference's ``VQCPCB`` class surface."""
