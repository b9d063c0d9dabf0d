"""Small shared helpers -- mirror of the reference's utils.py (exe_time :11-20, find_class_by_name :23-26,
get_local_time :29-31, get_latest_folder :34-52)."""
import datetime
import os
import time


def exe_time(func):
  """Decorator printing the wall time of `func` (utils.py:11-20)."""
  def timed(*args, **kwargs):
    t0 = time.time()
    back = func(*args, **kwargs)
    print("@%.3fs taken for {%s}" % (time.time() - t0, func.__name__))
    return back
  return timed


def find_class_by_name(name, modules):
  """First attribute called `name` among `modules`; StopIteration if none has it (utils.py:23-26)."""
  return next(a for a in (getattr(module, name, None) for module in modules) if a)


def get_local_time():
  return time.strftime("%y%m%d_%H%M%S", time.localtime())


def get_latest_folder(checkpoints_dir, nst_latest=1):
  """nst_latest-newest sub-folder by mtime, or checkpoints_dir itself when there are not enough (utils.py:34-52)."""
  folders = [os.path.join(checkpoints_dir, f) for f in os.listdir(checkpoints_dir)]
  folders = sorted((f for f in folders if os.path.isdir(f)), key=os.path.getmtime)
  if nst_latest < 1 or len(folders) < nst_latest:
    print("no %d-th newest folder under %s; returning it unchanged" % (nst_latest, checkpoints_dir))
    return checkpoints_dir
  return folders[-nst_latest]


def clean_file_by_time(log_dir, keepdays=7):
  """Delete files older than `keepdays` under log_dir (utils.py:55-67; the reference forgets to import datetime)."""
  limit = time.mktime((datetime.datetime.now() - datetime.timedelta(days=keepdays)).timetuple())
  for parent, _, filenames in os.walk(log_dir):
    for filename in filenames:
      full = os.path.join(parent, filename)
      if int(os.path.getctime(full)) < int(limit):
        os.remove(full)
