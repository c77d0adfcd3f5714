"""
Logger with the reference's levels and line format (tobias/utils/logger.py:27-205) so that tools parsing TOBIAS logs
(e.g. `Log2Table`) keep working: levels 0 silent, 1 warnings/errors (+ "comment"), 2 info, 3 STATS, 4 debug, 5 SPAM;
format "%(asctime)s (%(process)d) [%(levelname)s]\t%(message)s".  The GPU drivers are single-process, so the reference's
queue listener process (logger.py:124-160) has no counterpart here.
"""
import logging
import os
import sys
from datetime import datetime

from . import REFERENCE_VERSION, __version__


def add_logger_args(args):
    args.add_argument('--verbosity', metavar="<int>", help="Level of output logging (0: silent, 1: errors/warnings, 2: info, 3: stats, 4: debug, 5: spam) (default: 3)", choices=[0, 1, 2, 3, 4, 5], default=3, type=int)
    return args


class _Formatter(logging.Formatter):
    default_fmt = logging.Formatter("%(asctime)s (%(process)d) [%(levelname)s]\t%(message)s", "%Y-%m-%d %H:%M:%S")
    comment_fmt = logging.Formatter("%(message)s")

    def format(self, record):
        if record.levelname == "comment":
            return self.comment_fmt.format(record)
        return self.default_fmt.format(record)


class TobiasLogger(logging.Logger):
    logger_levels = {0: 0, 1: logging.WARNING, 2: logging.INFO, 3: int((logging.INFO + logging.DEBUG) / 2), 4: logging.DEBUG,
                     5: logging.DEBUG - 5}

    def __init__(self, tool_name="TOBIAS", level=3, queue=None):
        self.tool_name = tool_name
        logging.Logger.__init__(self, self.tool_name)
        if level == 0:
            self.disabled = True
        self._comment_level = self.logger_levels[1] + 1
        self._stats_level = self.logger_levels[3]
        self._spam_level = self.logger_levels[5]
        logging.addLevelName(self._comment_level, "comment")
        logging.addLevelName(self._stats_level, "STATS")
        logging.addLevelName(self._spam_level, "SPAM")
        self.setLevel(self.logger_levels[level])
        con = logging.StreamHandler(sys.stdout)
        con.setLevel(self.logger_levels[level])
        con.setFormatter(_Formatter())
        self.addHandler(con)
        self.begin_time = datetime.now()

    def comment(self, *args):
        self.log(self._comment_level, *args)

    def stats(self, *args):
        self.log(self._stats_level, *args)

    def spam(self, *args):
        self.log(self._spam_level, *args)

    def begin(self):
        self.cmd = "TOBIAS " + " ".join(sys.argv[1:])
        self.comment("# TOBIAS {0} {1} (run started {2})".format(REFERENCE_VERSION, self.tool_name, self.begin_time))
        self.comment("# tobias_b200 {0}: B200 (sm_100a) implementation of the per-base hot path".format(__version__))
        self.comment("# Working directory: {0}".format(os.getcwd()))
        self.comment("# Command line call: {0}\n".format(self.cmd))

    def end(self):
        total = datetime.now() - self.begin_time
        self.comment("")
        self.info("Finished {0} run (total time elapsed: {1})".format(self.tool_name, total))

    def arguments_overview(self, parser, args):
        content = "# ----- Input parameters -----\n"
        for group in parser._action_groups:
            for option in group._group_actions:
                if option.help != "==SUPPRESS==":
                    content += "# {0}:\t{1}\n".format(option.dest, getattr(args, option.dest, None))
        self.comment(content + "\n")

    def output_files(self, outfiles):
        self.comment("# ----- Output files -----")
        for outf in outfiles:
            if outf is not None:
                self.comment("# {0}".format(outf))
        self.comment("\n")

    # the reference drivers start / stop a listener process around the pools; nothing to do in a single process
    def start_logger_queue(self):
        self.queue = None

    def stop_logger_queue(self):
        pass
