"""pyflwdir_b200 -- B200-native D8 flow-network hot path behind pyflwdir's from_array / FlwdirRaster API."""
