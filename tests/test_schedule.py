"""The reference's only golden vector: the schedule known-answer test
(/root/reference test/runtests.jl:8-54), ported verbatim."""


def test_schedule_kat(pkg):
    t, f, N = True, False, None
    sched = pkg.MCMCSchedule(
        10, 4, [(1, range(3, 9)), (2, range(4, 11, 2)), (4, range(6, 11))],
        start=pkg.Step(N, N, 1, 1, (t, t, t)),
        extra_info=dict(layout_types=((1, 1, 1), (1, 2, 2), (1, 1, 3), (1, 2, 3)),
                        no_blocking=False, blocking_layout="not_needed"))
    expected = (
        (N, N, 1, 1, (t, t, t)), (1, 1, 1, 2, (t, f, f)), (1, 2, 1, 3, (t, f, f)), (1, 3, 1, 4, (t, f, t)),
        (1, 4, 2, 1, (t, f, f)), (2, 1, 2, 2, (t, f, f)), (2, 2, 2, 3, (t, f, f)), (2, 3, 2, 4, (t, f, t)),
        (2, 4, 3, 2, (t, t, f)), (3, 2, 3, 3, (t, f, f)), (3, 3, 3, 4, (t, f, t)),
        (3, 4, 4, 3, (t, f, t)), (4, 3, 4, 4, (t, f, t)),
        (4, 4, 5, 2, (t, t, f)), (5, 2, 5, 3, (t, f, f)), (5, 3, 5, 4, (t, f, t)),
        (5, 4, 6, 3, (t, f, t)),
        (6, 3, 7, 2, (t, f, f)), (7, 2, 7, 3, (t, f, f)),
        (7, 3, 8, 3, (t, t, t)),
        (8, 3, 9, 1, (t, t, f)), (9, 1, 9, 2, (t, f, f)), (9, 2, 9, 3, (t, f, f)),
        (9, 3, 10, 1, (t, t, f)), (10, 1, 10, 3, (t, t, f)),
    )
    got = [tuple(s) for s in sched]
    assert len(got) == len(expected) == 25
    for g, e in zip(got, expected):
        assert g == e


def test_init_block_layout(pkg):
    # src/schedule.jl:97-100: ((1, 1:n1), (2, 1:n2), ...)
    lay = pkg.init_block_layout([3, 5])
    assert lay == ((1, range(1, 4)), (2, range(1, 6)))
