def _noop(*a, **k):
    return None


clf = hist = savefig = close = scatter = plot = ylabel = title = show = figure = _noop
