"""Mirror of the reference's ``src.models`` package: the engine seam (``stage``) and the scorer
seam (``predictor``) that pipeline.py:14-15 imports, plus the Qwen2 shapes they run."""
