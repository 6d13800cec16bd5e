"""tnsu.utils: `l2` (the convergence distance of the run loop, reference src/tnsu/utils.py:78-85) and the plotting
entry point's NAME.  Plotting is out of scope of the engine: `plot_convergence_curve` draws
the logger with matplotlib when it is installed and says so when it is not."""
import numpy as np


def l2(x, y):
    """Euclidean distance between two weight vectors."""
    return np.sqrt(np.sum(np.square(x - y)))


def plot_convergence_curve(simple_update_object, figure_name="su_simulation_plot", figure_size=(17, 10),
                           floating_point_error=7, save_figure=False):
    try:
        import matplotlib.pyplot as plt
    except ImportError as exc:  # not installed in the build / GPU images
        raise ImportError("plot_convergence_curve needs matplotlib (plotting is not part of the B200 engine)") from exc
    log = simple_update_object.logger
    fig, ax = plt.subplots(figsize=figure_size)
    ax.semilogy(log["iteration"], log["error"], label="convergence error")
    ax.set_xlabel("Simple Update sweep")
    ax.set_ylabel("max_e ||lambda_e(t+1) - lambda_e(t)||")
    if log["energy"]:
        ax2 = ax.twinx()
        ax2.plot(log["iteration"], np.round(log["energy"], floating_point_error), color="tab:red", label="energy per site")
        ax2.set_ylabel("energy per site")
    ax.legend()
    if save_figure:
        fig.savefig(figure_name + ".png")
    return fig
