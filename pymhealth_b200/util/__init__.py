from .windows import rolling_apply, nonuniform_rolling_apply      # mirrors mhealth/util/__init__.py:1
