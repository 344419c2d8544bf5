"""Host-side mirrors of the reference's lib/utils helpers that sit on the rendering path."""
