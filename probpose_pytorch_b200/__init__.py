"""probpose_pytorch_b200 -- B200-native heatmap hot path of ProbPose (encode / decode / OKS loss)."""
