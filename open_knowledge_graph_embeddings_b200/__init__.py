"""B200-native hot path of OpenKGE (samuelbroscheit/open_knowledge_graph_embeddings).

Drop-in behind the reference's scorer x embedder mixin API (``openkge/model.py``) and the batch /
eval loop contract of ``openkge/trainer.py``; every hot op runs in hand-written sm_100a CUDA
kernels reached through the C ABI of ``include/okge_b200.h``. There is no CPU fallback.
"""
__version__ = "0.1.0"
