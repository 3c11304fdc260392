"""B200-native silhouette rasterizer behind EasyHeC's render_mask operator."""
__version__ = "0.1.0"
