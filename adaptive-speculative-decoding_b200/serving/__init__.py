"""Mirror of the reference's ``src.serving`` package for the hot path (pipeline, cache manager,
real-model pipeline).  The FastAPI shell (server.py) is out of scope."""
