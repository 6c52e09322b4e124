class ProgBar(object):
    def __init__(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def stop(self):
        pass
