"""Mirror of the reference package `tts.core.codec` for the decode direction only."""
