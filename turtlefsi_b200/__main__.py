from .run_turtle import main

main()
