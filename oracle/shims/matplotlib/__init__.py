"""matplotlib is not installed; the reference imports pyplot at module import only
(utils/anomaly_detection_utils.py:7, hyperspace/utils.py:5).  Plotting is out of scope."""
rcParams = {}
