from torchrecsys_b200.embeddings.init_embeddings import ScaledEmbedding, ZeroEmbedding  # noqa: F401
