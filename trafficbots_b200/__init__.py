"""trafficbots_b200 -- B200-native implementation of the TrafficBots hot path (scene encoding + closed-loop rollout).

Public entry points:
  trafficbots_b200.pl_modules.waymo_motion.WaymoMotion   drop-in for the reference LightningModule's hot-path methods
  trafficbots_b200.models.traffic_bots.TrafficBots       world-model shell (`encode_input_features`, `init`, `hidden`)
  trafficbots_b200.engine.Engine                         thin driver of the C ABI (include/trafficbots_b200.h)
"""
__version__ = "0.1.0"
