"""Host-side plumbing of the B200 pgm-vae hot path: ctypes binding (_ffi), data ingest (data)
and the data-parallel communicator (dist).  The reference-facing API lives in ``core``."""
