"""Mirror of ``PyContact.core`` for the contact-analysis hot path (GPU-backed)."""
