"""Host-side mirror of the reference's `mvdfusion` package (module paths match its yaml `target:` strings)."""
