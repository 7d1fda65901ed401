"""Device versions of the reference's plugins that follow `inference` in its README pipeline."""
