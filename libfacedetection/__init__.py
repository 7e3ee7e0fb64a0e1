# namespace holder for ``libfacedetection.train_b200`` (the B200-native YuNet hot path)
