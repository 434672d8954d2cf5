"""placeholder (filled in below)"""
timm_create_model = None
timm_create_optimizer_v2 = None
